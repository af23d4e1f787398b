//! World with the reference's public API (src/world.rs:11-153); `step` runs on the GPU through the C ABI.
use crate::arbiter::{Arbiter, ArbiterKey, ContactInfo, Edges, FeaturePair};
use crate::body::{Body, Shape};
use crate::errors::Sylt2DErrors;
use crate::ffi::*;
use crate::joint::Joint;
use crate::math_utils::{Mat2x2, MathErrors, Vec2};
use std::cell::{Ref, RefCell};
use std::collections::HashMap;
use std::rc::Rc;
use std::slice::Iter;

#[derive(Clone, Copy)]
pub struct WorldContext { pub accumulate_impulse: bool, pub warm_starting: bool, pub position_correction: bool }

pub struct World {
    #[allow(dead_code)] gravity: Vec2,
    #[allow(dead_code)] iterations: u32,
    pub world_context: WorldContext,
    pub bodies: Vec<Rc<RefCell<Body>>>,
    pub joints: Vec<Joint>,
    pub arbiters: HashMap<ArbiterKey, Arbiter>,
    handle: *mut SyltWorld,
    mirror: Vec<SyltBodyState>, // device state as of the last mirror, to detect caller-side mutation of pub fields
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
fn differs(a: &SyltBodyState, b: &SyltBodyState) -> bool {
    a.x.to_bits() != b.x.to_bits() || a.y.to_bits() != b.y.to_bits() || a.rotation.to_bits() != b.rotation.to_bits()
        || a.vx.to_bits() != b.vx.to_bits() || a.vy.to_bits() != b.vy.to_bits()
        || a.angular_velocity.to_bits() != b.angular_velocity.to_bits()
        || a.fx.to_bits() != b.fx.to_bits() || a.fy.to_bits() != b.fy.to_bits() || a.torque.to_bits() != b.torque.to_bits()
}

impl World {
    pub fn new(gravity: Vec2, iterations: u32) -> Self {
        let device = std::env::var("SYLT_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { sylt_world_create(gravity.x, gravity.y, iterations, device, &mut handle) };
        assert!(rc == SYLT_OK, "sylt_world_create failed ({}): no CUDA device? there is no CPU fallback", rc);
        Self {
            gravity, iterations,
            world_context: WorldContext { accumulate_impulse: true, warm_starting: false, position_correction: true },
            bodies: Vec::with_capacity(2), joints: Vec::with_capacity(2), arbiters: HashMap::new(), handle, mirror: Vec::new(),
        }
    }

    pub fn add_body(&mut self, body: Body) {
        let mut idx = 0i32;
        let rc = unsafe {
            match body.shape {
                Shape::Box => sylt_world_add_box(self.handle, body.id as u64, body.width.x, body.width.y, body.mass, body.friction,
                                                 body.position.x, body.position.y, body.rotation, body.velocity.x, body.velocity.y,
                                                 body.angular_velocity, &mut idx),
                Shape::ConvexPolygon => {
                    let flat: Vec<f32> = body.vertices.iter().flat_map(|v| [v.x, v.y]).collect();
                    sylt_world_add_polygon(self.handle, body.id as u64, flat.as_ptr(), body.vertices.len() as i32, body.mass, body.friction,
                                           body.position.x, body.position.y, body.rotation, body.velocity.x, body.velocity.y,
                                           body.angular_velocity, &mut idx)
                }
            }
        };
        assert!(rc == SYLT_OK, "add_body failed ({})", rc);
        let mut st = state_of(&body);
        st.fx = 0.0; st.fy = 0.0; st.torque = 0.0; // force/torque set before add_body are uploaded by the first step's diff
        self.mirror.push(st);
        self.bodies.push(Rc::new(RefCell::new(body)));
    }

    pub fn iter_bodies(&self) -> BodiesIter { BodiesIter { inner: self.bodies.iter() } }

    pub fn add_joint(&mut self, mut joint: Joint) {
        let mut idx = 0i32;
        let (id1, id2) = (joint.body_1.borrow().id as u64, joint.body_2.borrow().id as u64);
        let rc = unsafe { sylt_world_add_joint(self.handle, id1, id2, joint.anchor.x, joint.anchor.y, &mut idx) };
        assert!(rc == SYLT_OK, "couldn't find body in world bodies. ({})", rc);
        joint.index = Some(idx);
        self.joints.push(joint);
    }

    pub fn clear(&mut self) {
        unsafe { sylt_world_clear(self.handle) };
        self.bodies.clear(); self.joints.clear(); self.arbiters.clear(); self.mirror.clear();
    }

    fn upload_dirty(&mut self) {
        // callers mutate the pub fields between steps (e.g. examples/samples/src/main.rs:56-64): re-upload what changed
        let (mut first, mut run): (usize, Vec<SyltBodyState>) = (0, Vec::new());
        for (i, b) in self.bodies.iter().enumerate() {
            let s = state_of(&b.borrow());
            if differs(&s, &self.mirror[i]) {
                if run.is_empty() { first = i; }
                if first + run.len() != i { unsafe { sylt_world_set_bodies(self.handle, first as i32, run.len() as i32, run.as_ptr()) }; run.clear(); first = i; }
                run.push(s);
                self.mirror[i] = s;
            }
        }
        if !run.is_empty() { unsafe { sylt_world_set_bodies(self.handle, first as i32, run.len() as i32, run.as_ptr()) }; }
        for j in &self.joints {
            if let Some(k) = j.index { unsafe { sylt_world_set_joint_params(self.handle, k, j.softness, j.bias_factor) }; }
        }
        let c = self.world_context;
        unsafe { sylt_world_set_context(self.handle, c.accumulate_impulse as i32, c.warm_starting as i32, c.position_correction as i32) };
    }

    fn mirror_back(&mut self) {
        let n = self.bodies.len();
        self.mirror.resize(n, SyltBodyState::default());
        unsafe { sylt_world_get_bodies(self.handle, self.mirror.as_mut_ptr(), n as i32) };
        for (b, s) in self.bodies.iter().zip(self.mirror.iter()) {
            let mut b = b.borrow_mut();
            b.position = Vec2::new(s.x, s.y); b.rotation = s.rotation; b.velocity = Vec2::new(s.vx, s.vy);
            b.angular_velocity = s.angular_velocity; b.force = Vec2::new(s.fx, s.fy); b.torque = s.torque;
        }
        // arbiters: (ArbiterKey -> Arbiter) read back for the renderer (examples/samples/src/main.rs:550-570)
        let (na, nc) = unsafe { (sylt_world_num_arbiters(self.handle), sylt_world_num_contacts(self.handle)) };
        let mut arb = vec![SyltArbiterInfo::default(); na.max(0) as usize];
        let mut con = vec![SyltContact::default(); nc.max(0) as usize];
        let got = unsafe { sylt_world_get_arbiters(self.handle, arb.as_mut_ptr(), na, con.as_mut_ptr(), nc) };
        self.arbiters.clear();
        for a in arb.iter().take(got.max(0) as usize) {
            let cs = &con[a.contact_offset as usize..(a.contact_offset + a.num_contacts) as usize];
            let contacts = cs.iter().map(|c| Some(ContactInfo {
                position: Vec2::new(c.px, c.py), normal: Vec2::new(c.nx, c.ny), r1: Vec2::new(c.r1x, c.r1y), r2: Vec2::new(c.r2x, c.r2y),
                separation: c.separation, pn: c.pn, pt: c.pt, pnb: c.pnb, mass_normal: c.mass_normal, mass_tangent: c.mass_tangent, bias: c.bias,
                feature: FeaturePair { edges: Edges { in_edge_1: c.in_edge_1.into(), out_edge_1: c.out_edge_1.into(),
                                                      in_edge_2: c.in_edge_2.into(), out_edge_2: c.out_edge_2.into() }, value: c.feature_value },
            })).collect();
            self.arbiters.insert(ArbiterKey { body1_id: a.id1 as usize, body2_id: a.id2 as usize },
                                 Arbiter { friction: a.friction, num_contacts: a.num_contacts, contacts });
        }
    }

    fn map_err(&mut self, rc: i32) -> Result<(), Sylt2DErrors> {
        match rc {
            SYLT_OK => Ok(()),
            SYLT_ERR_NO_INVERSE => {
                let mut m = [0f32; 4];
                unsafe { sylt_world_last_error_matrix(self.handle, m.as_mut_ptr()) };
                Err(Sylt2DErrors::MathOperations(MathErrors::NoInverse { matrix: Mat2x2::new(Vec2::new(m[0], m[2]), Vec2::new(m[1], m[3])) }))
            }
            SYLT_ERR_BODY_ORDER => panic!("already borrowed: bodies must be added in increasing id order (reference: arbiter.rs:120-122)"),
            other => panic!("sylt2d_b200 backend error {}", other),
        }
    }

    pub fn broad_phase(&mut self) -> Result<(), Sylt2DErrors> {
        self.upload_dirty();
        let rc = unsafe { sylt_world_broad_phase(self.handle) };
        self.mirror_back();
        self.map_err(rc)
    }

    pub fn step(&mut self, dt: f32) -> Result<(), Sylt2DErrors> {
        self.upload_dirty();
        let rc = unsafe { sylt_world_step(self.handle, dt) };
        self.mirror_back();
        self.map_err(rc)
    }
}

impl Drop for World {
    fn drop(&mut self) { unsafe { sylt_world_destroy(self.handle) }; }
}
