//! Body: the same pub fields and constructors as the reference (src/body.rs:182-294).  Mass properties are computed
//! by the library from (shape, width / vertices, mass); the values mirrored here follow the same formulas so that
//! callers reading `inv_mass`, `moi`, ... see what the reference shows.
use crate::math_utils::Vec2;
use std::sync::atomic::{AtomicUsize, Ordering};

#[derive(Debug, Default, Clone, Copy, PartialEq)]
pub enum Shape { #[default] Box, ConvexPolygon }

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
        let (mut lo, mut hi) = (Vec2::new(f32::MAX, f32::MAX), Vec2::new(f32::MIN, f32::MIN));
        for v in &vertices {
            if v.x < lo.x { lo.x = v.x } if v.x > hi.x { hi.x = v.x } if v.y < lo.y { lo.y = v.y } if v.y > hi.y { hi.y = v.y }
        }
        Self {
            id: BODY_ID_COUNTER.fetch_add(1, Ordering::Relaxed), width: Vec2::new(hi.x - lo.x, hi.y - lo.y), mass, inv_mass, moi, inv_moi,
            vertices, shape: Shape::ConvexPolygon, ..Default::default()
        }
    }
    pub fn add_force(&mut self, force: Vec2) { self.force = self.force + force; }
    pub fn get_vertices(&self) -> Vec<Vec2> { self.vertices.clone() }
}
