//! World with the public API of the crate being replaced (its src/world.rs:11-153); `step` runs on the GPU through the C ABI.
//!
//! Per step: (1) `upload_dirty` — every pub field a caller may have written since the last mirror goes to the device (body state,
//! body properties, joint parameters and local anchors, the three context flags); (2) `sylt_world_step`; (3) `mirror_back` — body
//! state (36 B per body) into the `Rc<RefCell<Body>>`s; the arbiters stay on the GPU until `world.arbiters` is read (ArbiterMap).
//! Backend failures are never dropped: anything but Ok / NoInverse is a panic with the code, like the reference's own panics.
use crate::arbiter::{contact_from_ffi, Arbiter, ArbiterKey, ArbiterMap};
use crate::body::{Body, Shape};
use crate::errors::Sylt2DErrors;
use crate::ffi::*;
use crate::joint::Joint;
use crate::math_utils::{Mat2x2, MathErrors, Vec2};
use std::cell::{OnceCell, Ref, RefCell};
use std::collections::HashMap;
use std::rc::Rc;
use std::slice::Iter;

#[derive(Clone, Copy)]
pub struct WorldContext { pub accumulate_impulse: bool, pub warm_starting: bool, pub position_correction: bool }

/// The device handle, shared with the lazy arbiter map (which outlives no World: both are dropped together).
struct Handle(*mut SyltWorld);
impl Drop for Handle { fn drop(&mut self) { unsafe { sylt_world_destroy(self.0) }; } }

pub struct World {
    #[allow(dead_code)] gravity: Vec2,
    #[allow(dead_code)] iterations: u32,
    pub world_context: WorldContext,
    pub bodies: Vec<Rc<RefCell<Body>>>,
    pub joints: Vec<Joint>,
    pub arbiters: ArbiterMap,
    handle: Rc<Handle>,
    mirror: Vec<SyltBodyState>, // device state as of the last mirror, to detect caller-side mutation of pub fields
    props: Vec<[f32; 8]>,       // mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape as last sent
}

pub struct BodiesIter<'a> { inner: Iter<'a, Rc<RefCell<Body>>> }
impl<'a> Iterator for BodiesIter<'a> {
    type Item = Ref<'a, Body>;
    fn next(&mut self) -> Option<Self::Item> { self.inner.next().map(|b| b.borrow()) }
}

fn state_of(b: &Body) -> SyltBodyState {
    SyltBodyState { x: b.position.x, y: b.position.y, rotation: b.rotation, vx: b.velocity.x, vy: b.velocity.y,
                    angular_velocity: b.angular_velocity, fx: b.force.x, fy: b.force.y, torque: b.torque }
}
fn props_of(b: &Body) -> [f32; 8] {
    [b.mass, b.inv_mass, b.moi, b.inv_moi, b.width.x, b.width.y, b.friction, if b.shape == Shape::Box { 0.0 } else { 1.0 }]
}
fn bits_differ(a: &[f32], b: &[f32]) -> bool { a.iter().zip(b).any(|(x, y)| x.to_bits() != y.to_bits()) }
fn state_words(s: &SyltBodyState) -> [f32; 9] { [s.x, s.y, s.rotation, s.vx, s.vy, s.angular_velocity, s.fx, s.fy, s.torque] }

fn backend(rc: i32, what: &str) {
    if rc != SYLT_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(sylt_last_error_string()) }.to_string_lossy().into_owned();
        panic!("sylt2d_b200: {} failed with code {} {}", what, rc, msg);
    }
}

fn fetch_arbiters(h: *mut SyltWorld) -> HashMap<ArbiterKey, Arbiter> {
    let (na, nc) = unsafe { (sylt_world_num_arbiters(h), sylt_world_num_contacts(h)) };
    let mut arb = vec![SyltArbiterInfo::default(); na.max(0) as usize];
    let mut con = vec![SyltContact::default(); nc.max(0) as usize];
    let got = unsafe { sylt_world_get_arbiters(h, arb.as_mut_ptr(), na, con.as_mut_ptr(), nc) };
    assert!(got >= 0, "sylt2d_b200: sylt_world_get_arbiters failed with code {}", -got);
    let mut map = HashMap::with_capacity(got as usize);
    for a in arb.iter().take(got as usize) {
        let cs = &con[a.contact_offset as usize..(a.contact_offset + a.num_contacts) as usize];
        map.insert(ArbiterKey { body1_id: a.id1 as usize, body2_id: a.id2 as usize },
                   Arbiter { friction: a.friction, num_contacts: a.num_contacts, contacts: cs.iter().map(|c| Some(contact_from_ffi(c))).collect() });
    }
    map
}

impl World {
    pub fn new(gravity: Vec2, iterations: u32) -> Self {
        let device = std::env::var("SYLT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sylt_world_create(gravity.x, gravity.y, iterations, device, &mut raw) };
        assert!(rc == SYLT_OK, "sylt_world_create failed ({}): no CUDA device? there is no CPU fallback", rc);
        let handle = Rc::new(Handle(raw));
        let for_map = handle.clone();
        Self {
            gravity, iterations,
            world_context: WorldContext { accumulate_impulse: true, warm_starting: false, position_correction: true },
            bodies: Vec::with_capacity(2), joints: Vec::with_capacity(2),
            arbiters: ArbiterMap { cell: OnceCell::new(), fetch: Box::new(move || fetch_arbiters(for_map.0)) },
            handle, mirror: Vec::new(), props: Vec::new(),
        }
    }

    fn h(&self) -> *mut SyltWorld { self.handle.0 }

    pub fn add_body(&mut self, body: Body) {
        let mut idx = 0i32;
        let rc = unsafe {
            match body.shape {
                Shape::Box => sylt_world_add_box(self.h(), body.id as u64, body.width.x, body.width.y, body.mass, body.friction,
                                                 body.position.x, body.position.y, body.rotation, body.velocity.x, body.velocity.y,
                                                 body.angular_velocity, &mut idx),
                Shape::ConvexPolygon => {
                    let verts = body.get_polygon().get_vertices();
                    let flat: Vec<f32> = verts.iter().flat_map(|v| [v.x, v.y]).collect();
                    sylt_world_add_polygon(self.h(), body.id as u64, flat.as_ptr(), verts.len() as i32, body.mass, body.friction,
                                           body.position.x, body.position.y, body.rotation, body.velocity.x, body.velocity.y,
                                           body.angular_velocity, &mut idx)
                }
            }
        };
        backend(rc, "add_body");
        let mut st = state_of(&body);
        st.fx = 0.0; st.fy = 0.0; st.torque = 0.0; // force / torque set before add_body reach the device with the first step's diff
        self.mirror.push(st);
        // what the library derived from (shape, width, mass): if the caller changed inv_mass & co. by hand before add_body, the first diff sends them
        self.props.push(body.derived_props());
        self.bodies.push(Rc::new(RefCell::new(body)));
    }

    pub fn iter_bodies(&self) -> BodiesIter { BodiesIter { inner: self.bodies.iter() } }

    pub fn add_joint(&mut self, mut joint: Joint) {
        let mut idx = 0i32;
        let (id1, id2) = (joint.body_1.borrow().id as u64, joint.body_2.borrow().id as u64);
        // the local anchors as Joint::new computed them (or as the caller left them): nothing is re-derived from a world anchor
        let rc = unsafe {
            sylt_world_add_joint_local(self.h(), id1, id2, joint.local_anchor_1.x, joint.local_anchor_1.y, joint.local_anchor_2.x,
                                       joint.local_anchor_2.y, joint.softness, joint.bias_factor, &mut idx)
        };
        assert!(rc != SYLT_ERR_BODY_NOT_FOUND, "couldn't find body in world bodies.");
        backend(rc, "add_joint");
        joint.index = Some(idx);
        self.joints.push(joint);
    }

    pub fn clear(&mut self) {
        backend(unsafe { sylt_world_clear(self.h()) }, "clear");
        self.bodies.clear(); self.joints.clear(); self.arbiters.invalidate(); self.mirror.clear(); self.props.clear();
    }

    fn upload_dirty(&mut self) {
        // callers mutate the pub fields between steps (e.g. examples/samples/src/main.rs:56-64): runs of changed bodies go up in one call each
        let h = self.h();
        let (mut first, mut run): (usize, Vec<SyltBodyState>) = (0, Vec::new());
        let (mut pfirst, mut prun): (usize, Vec<f32>) = (0, Vec::new());
        for (i, b) in self.bodies.iter().enumerate() {
            let b = b.borrow();
            let s = state_of(&b);
            if bits_differ(&state_words(&s), &state_words(&self.mirror[i])) {
                if !run.is_empty() && first + run.len() != i {
                    backend(unsafe { sylt_world_set_bodies(h, first as i32, run.len() as i32, run.as_ptr()) }, "set_bodies");
                    run.clear();
                }
                if run.is_empty() { first = i; }
                run.push(s);
                self.mirror[i] = s;
            }
            let p = props_of(&b);
            if bits_differ(&p[..7], &self.props[i][..7]) {
                if !prun.is_empty() && pfirst + prun.len() / 8 != i {
                    backend(unsafe { sylt_world_set_body_props(h, pfirst as i32, (prun.len() / 8) as i32, prun.as_ptr()) }, "set_body_props");
                    prun.clear();
                }
                if prun.is_empty() { pfirst = i; }
                prun.extend_from_slice(&p);
                self.props[i] = p;
            }
        }
        if !run.is_empty() { backend(unsafe { sylt_world_set_bodies(h, first as i32, run.len() as i32, run.as_ptr()) }, "set_bodies"); }
        if !prun.is_empty() { backend(unsafe { sylt_world_set_body_props(h, pfirst as i32, (prun.len() / 8) as i32, prun.as_ptr()) }, "set_body_props"); }
        for j in &self.joints {  // (the library ignores values that did not change)
            if let Some(k) = j.index {
                backend(unsafe { sylt_world_set_joint_params(h, k, j.softness, j.bias_factor) }, "set_joint_params");
                backend(unsafe { sylt_world_set_joint_local_anchors(h, k, j.local_anchor_1.x, j.local_anchor_1.y, j.local_anchor_2.x, j.local_anchor_2.y) },
                        "set_joint_local_anchors");
            }
        }
        let c = self.world_context;
        backend(unsafe { sylt_world_set_context(h, c.accumulate_impulse as i32, c.warm_starting as i32, c.position_correction as i32) }, "set_context");
    }

    fn mirror_back(&mut self) {
        let n = self.bodies.len();
        self.mirror.resize(n, SyltBodyState::default());
        let got = unsafe { sylt_world_get_bodies(self.h(), self.mirror.as_mut_ptr(), n as i32) };
        assert!(got == n as i32, "sylt2d_b200: sylt_world_get_bodies returned {} for {} bodies", got, n);
        for (b, s) in self.bodies.iter().zip(self.mirror.iter()) {
            let mut b = b.borrow_mut();
            b.position = Vec2::new(s.x, s.y); b.rotation = s.rotation; b.velocity = Vec2::new(s.vx, s.vy);
            b.angular_velocity = s.angular_velocity; b.force = Vec2::new(s.fx, s.fy); b.torque = s.torque;
        }
        self.arbiters.invalidate();  // (ArbiterKey -> Arbiter) is fetched when a caller reads it (examples/samples/src/main.rs:550-570)
    }

    fn map_err(&mut self, rc: i32) -> Result<(), Sylt2DErrors> {
        match rc {
            SYLT_OK => Ok(()),
            SYLT_ERR_NO_INVERSE => {
                let mut m = [0f32; 4];
                unsafe { sylt_world_last_error_matrix(self.h(), m.as_mut_ptr()) };
                Err(Sylt2DErrors::MathOperations(MathErrors::NoInverse { matrix: Mat2x2::new(Vec2::new(m[0], m[2]), Vec2::new(m[1], m[3])) }))
            }
            SYLT_ERR_BODY_ORDER => panic!("already borrowed: bodies must be added in increasing id order (reference: arbiter.rs:120-122)"),
            other => { backend(other, "step"); Ok(()) }
        }
    }

    pub fn broad_phase(&mut self) -> Result<(), Sylt2DErrors> {
        self.upload_dirty();
        let rc = unsafe { sylt_world_broad_phase(self.h()) };
        self.mirror_back();
        self.map_err(rc)
    }

    pub fn step(&mut self, dt: f32) -> Result<(), Sylt2DErrors> {
        self.upload_dirty();
        let rc = unsafe { sylt_world_step(self.h(), dt) };
        self.mirror_back();
        self.map_err(rc)
    }
}
