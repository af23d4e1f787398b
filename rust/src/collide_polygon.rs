//! `collide_polygons(&mut contacts, &b1, &b2) -> i32` — the convex-polygon narrow phase as a free function (any pair with at
//! least one polygon; a box takes part with the four vertices stored at `Body::new`).  One pair through `sylt_collide_pairs`.
use crate::arbiter::{contact_from_ffi, Contact};
use crate::body::Body;
use crate::ffi::{desc_of, sylt_collide_pairs, SyltContact, SYLT_MAX_MANIFOLD, SYLT_OK};

/// When the polygons intersect, `contacts` is REPLACED by the new manifold; the return value is `contacts.len()` either way
/// (so a miss reports whatever the vector already held) — both as in the function this replaces.
pub fn collide_polygons(contacts: &mut Vec<Contact>, b1: &Body, b2: &Body) -> i32 {
    let (v1, v2) = (b1.get_polygon().get_vertices(), b2.get_polygon().get_vertices());
    let flat: Vec<f32> = v1.iter().chain(v2.iter()).flat_map(|v| [v.x, v.y]).collect();
    // both bodies go in as polygons: the polygon path is what the reference runs for this call whatever the shapes are
    let (mut a, mut b) = (desc_of(b1, 0, v1.len() as i32), desc_of(b2, v1.len() as i32, v2.len() as i32));
    a.shape = crate::ffi::SYLT_SHAPE_POLYGON;
    b.shape = crate::ffi::SYLT_SHAPE_POLYGON;
    let mut out = vec![SyltContact::default(); SYLT_MAX_MANIFOLD];
    let mut n = 0i32;
    let device = std::env::var("SYLT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
    let rc = unsafe {
        sylt_collide_pairs(device, &a, &b, 1, flat.as_ptr(), (v1.len() + v2.len()) as i32, out.as_mut_ptr(), SYLT_MAX_MANIFOLD as i32, &mut n)
    };
    assert!(rc == SYLT_OK, "sylt_collide_pairs failed ({})", rc);
    if n > 0 { *contacts = out.iter().take(n as usize).map(|c| Some(contact_from_ffi(c))).collect(); }
    contacts.len() as i32
}
