//! Joint::new(body_1, body_2, anchor, &world) with pub softness / bias_factor / local anchors (reference: src/joint.rs:10-55).
//! All four are LIVE: World::step sends them before every step (the library ignores unchanged values).
use crate::body::Body;
use crate::math_utils::{Mat2x2, Vec2};
use crate::world::World;
use std::cell::RefCell;
use std::rc::Rc;

pub struct Joint {
    pub bias_factor: f32,
    pub softness: f32,
    pub local_anchor_1: Vec2,
    pub local_anchor_2: Vec2,
    pub body_1: Rc<RefCell<Body>>,
    pub body_2: Rc<RefCell<Body>>,
    pub(crate) index: Option<i32>,
}

impl Joint {
    pub fn new(body_1: Body, body_2: Body, anchor: Vec2, world: &World) -> Self {
        let b1 = world.bodies.iter().find(|b| b.borrow().id == body_1.id).expect("couldn't find body 1 in world bodies.");
        let b2 = world.bodies.iter().find(|b| b.borrow().id == body_2.id).expect("couldn't find body 2 in world bodies.");
        let rt1 = Mat2x2::new_from_angle(b1.borrow().rotation).transpose();
        let rt2 = Mat2x2::new_from_angle(b2.borrow().rotation).transpose();
        Self {
            body_1: b1.clone(), body_2: b2.clone(),
            local_anchor_1: rt1 * (anchor - b1.borrow().position),
            local_anchor_2: rt2 * (anchor - b2.borrow().position),
            softness: 0.0, bias_factor: 0.2, index: None,
        }
    }
}
