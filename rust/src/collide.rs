//! `collide(&mut contacts, &a, &b) -> i32` — the box-box narrow phase as a free function (callers: examples/collision-debug,
//! examples/simple_collide.rs).  The arithmetic runs on the GPU: one pair through `sylt_collide_pairs`.
use crate::arbiter::{contact_from_ffi, Contact};
use crate::body::Body;
use crate::ffi::{desc_of, sylt_collide_pairs, SyltContact, SYLT_OK};

/// Appends the contacts of the pair to `contacts` and returns how many were added (0..=2), like the function it replaces.
pub fn collide(contacts: &mut Vec<Contact>, body_a: &Body, body_b: &Body) -> i32 {
    // `collide` treats both bodies as boxes of `width`, whatever their shape field says
    let (mut a, mut b) = (desc_of(body_a, 0, 0), desc_of(body_b, 0, 0));
    a.shape = crate::ffi::SYLT_SHAPE_BOX;
    b.shape = crate::ffi::SYLT_SHAPE_BOX;
    let mut out = [SyltContact::default(); 2];
    let mut n = 0i32;
    let device = std::env::var("SYLT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let rc = unsafe { sylt_collide_pairs(device, &a, &b, 1, std::ptr::null(), 0, out.as_mut_ptr(), 2, &mut n) };
    assert!(rc == SYLT_OK, "sylt_collide_pairs failed ({}): no CUDA device? there is no CPU fallback", rc);
    for c in out.iter().take(n.clamp(0, 2) as usize) { contacts.push(Some(contact_from_ffi(c))); }
    n
}
