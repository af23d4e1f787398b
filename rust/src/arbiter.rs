//! Arbiter / ContactInfo as read-only mirrors of the device state (reference: src/arbiter.rs:27-114).
use crate::body::Body;
use crate::ffi::SyltContact;
use crate::math_utils::Vec2;
use std::cell::OnceCell;
use std::collections::HashMap;
use std::fmt;
use std::ops::Deref;

#[derive(Debug)]
pub enum ArbiterErrors { NoOldContactFound, NoNewContactFound }
impl fmt::Display for ArbiterErrors {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            ArbiterErrors::NoOldContactFound => write!(f, "No old contacts found."),
            ArbiterErrors::NoNewContactFound => write!(f, "No new contacts found."),
        }
    }
}
impl std::error::Error for ArbiterErrors {}

#[derive(Debug, Clone, Copy, Default, PartialEq)]
pub enum EdgeNumbers { #[default] NoEdge = 0, Edge1, Edge2, Edge3, Edge4 }
impl From<u8> for EdgeNumbers {
    fn from(v: u8) -> Self { match v { 1 => Self::Edge1, 2 => Self::Edge2, 3 => Self::Edge3, 4 => Self::Edge4, _ => Self::NoEdge } }
}
#[derive(Debug, Clone, Copy, Default)]
pub struct Edges { pub in_edge_1: EdgeNumbers, pub out_edge_1: EdgeNumbers, pub in_edge_2: EdgeNumbers, pub out_edge_2: EdgeNumbers }
#[derive(Debug, Default, Clone, Copy)]
pub struct FeaturePair { pub edges: Edges, pub value: i32 }
impl FeaturePair { pub fn new(edges: Edges, value: i32) -> Self { Self { edges, value } } }
pub type Contact = Option<ContactInfo>;

#[derive(Debug, Default, Clone, Copy)]
pub struct ContactInfo {
    pub position: Vec2, pub normal: Vec2, pub r1: Vec2, pub r2: Vec2,
    pub separation: f32, pub pn: f32, pub pt: f32, pub pnb: f32,
    pub mass_normal: f32, pub mass_tangent: f32, pub bias: f32,
    pub feature: FeaturePair,
}

#[derive(Debug, Eq, Hash, PartialEq, Clone, Copy)]
pub struct ArbiterKey { pub(crate) body1_id: usize, pub(crate) body2_id: usize }
impl ArbiterKey {
    /// (lower id, higher id), whatever the argument order
    pub fn new(body_1: &Body, body_2: &Body) -> Self {
        Self { body1_id: body_1.id.min(body_2.id), body2_id: body_1.id.max(body_2.id) }
    }
}

pub(crate) fn contact_from_ffi(c: &SyltContact) -> ContactInfo {
    ContactInfo {
        position: Vec2::new(c.px, c.py), normal: Vec2::new(c.nx, c.ny), r1: Vec2::new(c.r1x, c.r1y), r2: Vec2::new(c.r2x, c.r2y),
        separation: c.separation, pn: c.pn, pt: c.pt, pnb: c.pnb, mass_normal: c.mass_normal, mass_tangent: c.mass_tangent, bias: c.bias,
        feature: FeaturePair { edges: Edges { in_edge_1: c.in_edge_1.into(), out_edge_1: c.out_edge_1.into(),
                                              in_edge_2: c.in_edge_2.into(), out_edge_2: c.out_edge_2.into() }, value: c.feature_value },
    }
}

/// `World.arbiters`: reads like the `HashMap<ArbiterKey, Arbiter>` it stands in for (`.iter()`, `.get()`, `.len()`, `{:?}`,
/// `for (k, a) in &world.arbiters`) but is filled LAZILY: the contacts stay on the GPU until the first read after a step
/// (34 MB and 230 k allocations per step at 100 k bodies otherwise — a renderer pays that, a headless caller never does).
pub struct ArbiterMap {
    pub(crate) cell: OnceCell<HashMap<ArbiterKey, Arbiter>>,
    pub(crate) fetch: Box<dyn Fn() -> HashMap<ArbiterKey, Arbiter>>,
}
impl ArbiterMap {
    pub(crate) fn invalidate(&mut self) { self.cell = OnceCell::new(); }
}
impl Deref for ArbiterMap {
    type Target = HashMap<ArbiterKey, Arbiter>;
    fn deref(&self) -> &Self::Target { self.cell.get_or_init(|| (self.fetch)()) }
}
impl fmt::Debug for ArbiterMap {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { self.deref().fmt(f) }
}
impl<'a> IntoIterator for &'a ArbiterMap {
    type Item = (&'a ArbiterKey, &'a Arbiter);
    type IntoIter = std::collections::hash_map::Iter<'a, ArbiterKey, Arbiter>;
    fn into_iter(self) -> Self::IntoIter { self.deref().iter() }
}

#[derive(Debug)]
pub struct Arbiter {
    pub(crate) friction: f32,
    pub num_contacts: i32,
    pub contacts: Vec<Contact>,
}
