//! extern "C" declarations of include/sylt2d_b200.h (the subset the wrapper uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const SYLT_OK: c_int = 0;
pub const SYLT_ERR_NO_INVERSE: c_int = 1;
pub const SYLT_ERR_ARBITER: c_int = 2;
pub const SYLT_ERR_BODY_ORDER: c_int = 3;
pub const SYLT_ERR_BODY_NOT_FOUND: c_int = 4;
pub const SYLT_ERR_CAPACITY: c_int = 5;
pub const SYLT_ERR_CUDA: c_int = 6;
pub const SYLT_SHAPE_BOX: i32 = 0;
pub const SYLT_SHAPE_POLYGON: i32 = 1;
pub const SYLT_MAX_POLY_VERTS: usize = 32;
pub const SYLT_MAX_MANIFOLD: usize = 2 * SYLT_MAX_POLY_VERTS;

/// Body construction record (64 bytes, include/sylt2d_b200.h `SyltBodyDesc`)
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SyltBodyDesc {
    pub id: u64, pub shape: i32, pub vert_offset: i32, pub vert_count: i32, pub group: i32,
    pub width_x: f32, pub width_y: f32, pub mass: f32, pub friction: f32,
    pub x: f32, pub y: f32, pub rotation: f32, pub vx: f32, pub vy: f32, pub angular_velocity: f32,
}

pub fn desc_of(b: &crate::body::Body, vert_offset: i32, vert_count: i32) -> SyltBodyDesc {
    SyltBodyDesc {
        id: b.id as u64,
        shape: match b.shape { crate::body::Shape::Box => SYLT_SHAPE_BOX, crate::body::Shape::ConvexPolygon => SYLT_SHAPE_POLYGON },
        vert_offset, vert_count, group: 0, width_x: b.width.x, width_y: b.width.y, mass: b.mass, friction: b.friction,
        x: b.position.x, y: b.position.y, rotation: b.rotation, vx: b.velocity.x, vy: b.velocity.y, angular_velocity: b.angular_velocity,
    }
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SyltBodyState {
    pub x: f32, pub y: f32, pub rotation: f32, pub vx: f32, pub vy: f32, pub angular_velocity: f32,
    pub fx: f32, pub fy: f32, pub torque: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SyltContact {
    pub px: f32, pub py: f32, pub nx: f32, pub ny: f32, pub r1x: f32, pub r1y: f32, pub r2x: f32, pub r2y: f32,
    pub separation: f32, pub pn: f32, pub pt: f32, pub pnb: f32, pub mass_normal: f32, pub mass_tangent: f32, pub bias: f32,
    pub in_edge_1: u8, pub out_edge_1: u8, pub in_edge_2: u8, pub out_edge_2: u8,
    pub feature_value: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SyltArbiterInfo {
    pub id1: u64, pub id2: u64, pub body1: i32, pub body2: i32, pub friction: f32,
    pub num_contacts: i32, pub contact_offset: i32, pub color: i32,
}

pub enum SyltWorld {}

extern "C" {
    pub fn sylt_last_error_string() -> *const c_char;
    pub fn sylt_world_create(gx: f32, gy: f32, iterations: u32, device: c_int, out: *mut *mut SyltWorld) -> c_int;
    pub fn sylt_world_destroy(w: *mut SyltWorld) -> c_int;
    pub fn sylt_world_clear(w: *mut SyltWorld) -> c_int;
    pub fn sylt_world_set_context(w: *mut SyltWorld, accumulate: c_int, warm: c_int, poscorr: c_int) -> c_int;
    pub fn sylt_world_add_box(w: *mut SyltWorld, id: u64, wx: f32, wy: f32, mass: f32, friction: f32, x: f32, y: f32, rot: f32,
                              vx: f32, vy: f32, omega: f32, out_index: *mut i32) -> c_int;
    pub fn sylt_world_add_polygon(w: *mut SyltWorld, id: u64, verts_xy: *const f32, n: i32, mass: f32, friction: f32, x: f32, y: f32,
                                  rot: f32, vx: f32, vy: f32, omega: f32, out_index: *mut i32) -> c_int;
    pub fn sylt_world_add_joint(w: *mut SyltWorld, id1: u64, id2: u64, ax: f32, ay: f32, out_index: *mut i32) -> c_int;
    pub fn sylt_world_set_joint_params(w: *mut SyltWorld, joint: i32, softness: f32, bias_factor: f32) -> c_int;
    pub fn sylt_world_add_joint_local(w: *mut SyltWorld, id1: u64, id2: u64, la1x: f32, la1y: f32, la2x: f32, la2y: f32,
                                      softness: f32, bias_factor: f32, out_index: *mut i32) -> c_int;
    pub fn sylt_world_set_joint_local_anchors(w: *mut SyltWorld, joint: i32, la1x: f32, la1y: f32, la2x: f32, la2y: f32) -> c_int;
    pub fn sylt_world_get_body_props(w: *mut SyltWorld, out8: *mut f32, cap: i32) -> c_int;
    pub fn sylt_world_set_body_props(w: *mut SyltWorld, first: i32, count: i32, props8: *const f32) -> c_int;
    pub fn sylt_collide_pairs(device: c_int, a: *const SyltBodyDesc, b: *const SyltBodyDesc, n: i32, verts_xy: *const f32, n_verts: i32,
                              out: *mut SyltContact, cap: i32, out_counts: *mut i32) -> c_int;
    pub fn sylt_world_num_arbiters(w: *mut SyltWorld) -> c_int;
    pub fn sylt_world_num_contacts(w: *mut SyltWorld) -> c_int;
    pub fn sylt_world_get_bodies(w: *mut SyltWorld, out: *mut SyltBodyState, cap: i32) -> c_int;
    pub fn sylt_world_set_bodies(w: *mut SyltWorld, first: i32, count: i32, states: *const SyltBodyState) -> c_int;
    pub fn sylt_world_get_arbiters(w: *mut SyltWorld, arb: *mut SyltArbiterInfo, acap: i32, con: *mut SyltContact, ccap: i32) -> c_int;
    pub fn sylt_world_broad_phase(w: *mut SyltWorld) -> c_int;
    pub fn sylt_world_step(w: *mut SyltWorld, dt: f32) -> c_int;
    pub fn sylt_world_last_error_matrix(w: *mut SyltWorld, out4: *mut f32) -> c_int;
}
