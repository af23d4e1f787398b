//! Body: the same pub fields and constructors as the reference (src/body.rs:182-294).  Mass properties are computed
//! by the library from (shape, width / vertices, mass); the values mirrored here follow the same formulas so that
//! callers reading `inv_mass`, `moi`, ... see what the reference shows.
use crate::math_utils::{Mat2x2, Vec2};
use std::sync::atomic::{AtomicUsize, Ordering};

/// Host-side polygon helper with the public surface of the crate being replaced (callers build shapes with it and draw
/// `body.get_polygon().rotate(r).translate(p).get_vertices()`).  Quirks kept on purpose: `get_vertex(i)` is shifted by one
/// (cyclic, so geometry is unaffected), `area()` is unsigned.
pub struct ConvexPolygon { vertices: Vec<Vec2> }

impl ConvexPolygon {
    pub fn new(vertices: Vec<Vec2>) -> Self { Self { vertices } }
    pub fn get_num_vertices(&self) -> usize { self.vertices.len() }
    pub fn get_vertex(&self, i: isize) -> Vec2 {
        let n = self.vertices.len() as isize;
        self.vertices[((i + n + 1) % n) as usize]
    }
    pub fn get_edge(&self, i: isize) -> Vec2 { self.get_vertex(i + 1) - self.get_vertex(i) }
    pub fn get_normal(&self, i: isize) -> Vec2 { let e = self.get_edge(i); Vec2::new(e.y, -e.x) }
    pub fn area(&self) -> f32 {
        let twice: f32 = (0..self.vertices.len()).map(|i| shifted(&self.vertices, i).cross_z(shifted(&self.vertices, i + 1))).sum();
        twice.abs() / 2.0
    }
    pub fn centroid(&self) -> Vec2 { centroid(&self.vertices) }
    pub fn moi(&self) -> f32 { area_moment(&self.vertices) }
    pub fn scale(&mut self, factor: f32) {
        let c = self.centroid();
        for v in self.vertices.iter_mut() { *v = ((*v - c) * factor) + c; }
    }
    pub fn bounding_box(&self) -> Vec2 { bounding_box(&self.vertices) }
    pub fn rotate(&self, angle: f32) -> ConvexPolygon {
        let (c, r) = (self.centroid(), Mat2x2::new_from_angle(angle));
        ConvexPolygon { vertices: self.vertices.iter().map(|&v| r * Vec2::new(v.x - c.x, v.y - c.y)).collect() }
    }
    pub fn translate(&self, position: Vec2) -> ConvexPolygon {
        ConvexPolygon { vertices: self.vertices.iter().map(|&v| v + position).collect() }
    }
    pub fn get_polygon(&self) -> ConvexPolygon { ConvexPolygon { vertices: self.vertices.clone() } }

    /// (mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape) as the library derives them from (shape, width or
    /// vertices, mass) at add_body — the baseline the wrapper diffs the pub fields against.
    pub(crate) fn derived_props(&self) -> [f32; 8] {
        let (width, unit_moi) = match self.shape {
            Shape::Box => (self.width, (self.width.x * self.width.x + self.width.y * self.width.y) / 12.0),
            Shape::ConvexPolygon => (bounding_box(&self.vertices), area_moment(&self.vertices)),
        };
        let (inv_mass, moi, inv_moi) = if self.mass < f32::MAX {
            let moi = if self.shape == Shape::Box { self.mass * (self.width.x * self.width.x + self.width.y * self.width.y) / 12.0 } else { self.mass * unit_moi };
            (1.0 / self.mass, moi, 1.0 / moi)
        } else { (0.0, f32::MAX, 0.0) };
        [self.mass, inv_mass, moi, inv_moi, width.x, width.y, self.friction, if self.shape == Shape::Box { 0.0 } else { 1.0 }]
    }
}

trait CrossZ { fn cross_z(self, o: Vec2) -> f32; }
impl CrossZ for Vec2 { fn cross_z(self, o: Vec2) -> f32 { self.x * o.y - self.y * o.x } }

fn bounding_box(vs: &[Vec2]) -> Vec2 {
    let (mut lo, mut hi) = (Vec2::new(f32::MAX, f32::MAX), Vec2::new(f32::MIN, f32::MIN));
    for v in vs {
        if v.x < lo.x { lo.x = v.x } if v.x > hi.x { hi.x = v.x } if v.y < lo.y { lo.y = v.y } if v.y > hi.y { hi.y = v.y }
    }
    Vec2::new(hi.x - lo.x, hi.y - lo.y)
}

#[derive(Debug, Default, Clone, Copy, PartialEq)]
pub enum Shape { #[default] Box, ConvexPolygon }

/// Which pub fields are LIVE after `World::add_body`: position, rotation, velocity, angular_velocity, force, torque (uploaded
/// when they differ from the last mirror) and friction, mass, inv_mass, moi, inv_moi, width (`sylt_world_set_body_props`) — the
/// fields the replaced crate reads on every step.  `id` and `shape` are fixed at `add_body`; the vertices are private.

#[derive(Debug, Default, Clone)]
pub struct Body {
    pub id: usize,
    pub position: Vec2,
    pub rotation: f32,
    pub velocity: Vec2,
    pub angular_velocity: f32,
    pub force: Vec2,
    pub torque: f32,
    pub width: Vec2,
    pub friction: f32,
    pub mass: f32,
    pub inv_mass: f32,
    pub moi: f32,
    pub inv_moi: f32,
    pub(crate) vertices: Vec<Vec2>,
    pub shape: Shape,
}

static BODY_ID_COUNTER: AtomicUsize = AtomicUsize::new(1);

fn shifted(v: &[Vec2], i: usize) -> Vec2 { v[(i + 1) % v.len()] } // ConvexPolygon::get_vertex (+1 shift)

fn centroid(v: &[Vec2]) -> Vec2 {
    let (mut cx, mut cy, mut area) = (0.0f32, 0.0f32, 0.0f32);
    for i in 0..v.len() {
        let (p1, p2) = (shifted(v, i), shifted(v, i + 1));
        let c = p1.x * p2.y - p1.y * p2.x;
        area += c; cx += (p1.x + p2.x) * c; cy += (p1.y + p2.y) * c;
    }
    area /= 2.0; cx /= 6.0 * area; cy /= 6.0 * area;
    Vec2 { x: cx, y: cy }
}

fn area_moment(v: &[Vec2]) -> f32 {
    let c = centroid(v);
    let mut moi = 0.0f32;
    for i in 0..v.len() {
        let (a, b) = (shifted(v, i), shifted(v, i + 1));
        let (p1, p2) = (Vec2::new(a.x - c.x, a.y - c.y), Vec2::new(b.x - c.x, b.y - c.y));
        let cr = p1.x * p2.y - p1.y * p2.x;
        moi += cr * (p1.x * p1.x + p1.x * p2.x + p2.x * p2.x + p1.y * p1.y + p1.y * p2.y + p2.y * p2.y);
    }
    moi.abs() / 12.0
}

impl Body {
    pub fn new(width: Vec2, mass: f32) -> Self {
        let (inv_mass, moi, inv_moi) = if mass < f32::MAX {
            let moi = mass * (width.x * width.x + width.y * width.y) / 12.0;
            (1.0 / mass, moi, 1.0 / moi)
        } else { (0.0, f32::MAX, 0.0) };
        let (hw, hh) = (width.x / 2.0, width.y / 2.0);
        Self {
            id: BODY_ID_COUNTER.fetch_add(1, Ordering::Relaxed), width, mass, inv_mass, moi, inv_moi,
            vertices: vec![Vec2::new(hw, hh), Vec2::new(-hw, hh), Vec2::new(-hw, -hh), Vec2::new(hw, -hh)],
            shape: Shape::Box, ..Default::default()
        }
    }
    pub fn new_polygon(vertices: Vec<Vec2>, mass: f32) -> Self {
        let (inv_mass, moi, inv_moi) = if mass < f32::MAX {
            let moi = mass * area_moment(&vertices);
            (1.0 / mass, moi, 1.0 / moi)
        } else { (0.0, f32::MAX, 0.0) };
        Self {
            id: BODY_ID_COUNTER.fetch_add(1, Ordering::Relaxed), width: bounding_box(&vertices), mass, inv_mass, moi, inv_moi,
            vertices, shape: Shape::ConvexPolygon, ..Default::default()
        }
    }
    pub fn add_force(&mut self, force: Vec2) { self.force = self.force + force; }
    pub fn get_polygon(&self) -> ConvexPolygon { ConvexPolygon { vertices: self.vertices.clone() } }

    /// (mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape) as the library derives them from (shape, width or
    /// vertices, mass) at add_body — the baseline the wrapper diffs the pub fields against.
    pub(crate) fn derived_props(&self) -> [f32; 8] {
        let (width, unit_moi) = match self.shape {
            Shape::Box => (self.width, (self.width.x * self.width.x + self.width.y * self.width.y) / 12.0),
            Shape::ConvexPolygon => (bounding_box(&self.vertices), area_moment(&self.vertices)),
        };
        let (inv_mass, moi, inv_moi) = if self.mass < f32::MAX {
            let moi = if self.shape == Shape::Box { self.mass * (self.width.x * self.width.x + self.width.y * self.width.y) / 12.0 } else { self.mass * unit_moi };
            (1.0 / self.mass, moi, 1.0 / moi)
        } else { (0.0, f32::MAX, 0.0) };
        [self.mass, inv_mass, moi, inv_moi, width.x, width.y, self.friction, if self.shape == Shape::Box { 0.0 } else { 1.0 }]
    }
}
