//! Arbiter / ContactInfo as read-only mirrors of the device state (reference: src/arbiter.rs:27-114).
use crate::math_utils::Vec2;
use std::fmt;

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

#[derive(Debug)]
pub struct Arbiter {
    pub(crate) friction: f32,
    pub num_contacts: i32,
    pub contacts: Vec<Contact>,
}
