//! Vec2 / Mat2x2 as the reference exposes them (src/math_utils.rs:23-248).  Host-side only: the step arithmetic
//! runs on the GPU.  `Mat2x2::invert` keeps the reference's behaviour (adjugate * det), since callers may depend on it.
use std::fmt::{self, Display};
use std::ops::{Add, Mul, Neg, Sub};

#[derive(Debug)]
pub enum MathErrors {
    NoInverse { matrix: Mat2x2 },
}
impl Display for MathErrors {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            MathErrors::NoInverse { matrix } => write!(f, "No inverse could be found for matrix {}", matrix),
        }
    }
}
impl std::error::Error for MathErrors {}

#[derive(Debug, Default, PartialOrd, Clone, Copy)]
pub struct Vec2 { pub x: f32, pub y: f32 }

impl PartialEq for Vec2 {
    fn eq(&self, o: &Self) -> bool { (self.x - o.x).abs() < f32::EPSILON && (self.y - o.y).abs() < f32::EPSILON }
}
impl Vec2 {
    pub fn new(x: f32, y: f32) -> Vec2 { Self { x, y } }
    pub fn length(self) -> f32 { (self.x * self.x + self.y * self.y).sqrt() }
    pub fn dot(&self, r: Vec2) -> f32 { self.x * r.x + self.y * r.y }
    pub fn abs(&self) -> Self { Self { x: self.x.abs(), y: self.y.abs() } }
}
impl Display for Vec2 {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "({}, {})", self.x, self.y) }
}
pub trait Cross<T> { type Output; fn cross(&self, rhs: T) -> Self::Output; }
impl Cross<Vec2> for Vec2 { type Output = f32; fn cross(&self, r: Vec2) -> f32 { self.x * r.y - self.y * r.x } }
impl Cross<f32> for Vec2 { type Output = Vec2; fn cross(&self, r: f32) -> Vec2 { Vec2 { x: self.y * r, y: self.x * -r } } }
impl Cross<Vec2> for f32 { type Output = Vec2; fn cross(&self, r: Vec2) -> Vec2 { Vec2 { x: -self * r.y, y: self * r.x } } }
impl Add for Vec2 { type Output = Self; fn add(self, r: Self) -> Self { Self { x: self.x + r.x, y: self.y + r.y } } }
impl Sub for Vec2 { type Output = Self; fn sub(self, r: Self) -> Self { Self { x: self.x - r.x, y: self.y - r.y } } }
impl Mul<f32> for Vec2 { type Output = Self; fn mul(self, r: f32) -> Self { Self { x: self.x * r, y: self.y * r } } }
impl Neg for Vec2 { type Output = Vec2; fn neg(self) -> Vec2 { Vec2 { x: -self.x, y: -self.y } } }
impl From<Vec2> for (f32, f32) { fn from(v: Vec2) -> Self { (v.x, v.y) } }

#[derive(Debug, Default, Clone, Copy, PartialEq)]
pub struct Mat2x2 { pub col1: Vec2, pub col2: Vec2 }
impl Mat2x2 {
    pub fn new_from_angle(angle: f32) -> Self {
        let (c, s) = (f32::cos(angle), f32::sin(angle));
        Self { col1: Vec2::new(c, s), col2: Vec2::new(-s, c) }
    }
    pub fn new(col1: Vec2, col2: Vec2) -> Self { Self { col1, col2 } }
    pub fn transpose(&self) -> Self {
        Self { col1: Vec2::new(self.col1.x, self.col2.x), col2: Vec2::new(self.col1.y, self.col2.y) }
    }
    pub fn invert(&self) -> Result<Self, MathErrors> {
        let (a, b, c, d) = (self.col1.x, self.col2.x, self.col1.y, self.col2.y);
        let det = a * d - b * c;
        if det == 0.0 { Err(MathErrors::NoInverse { matrix: *self }) }
        else { Ok(Self { col1: Vec2::new(det * d, -det * c), col2: Vec2::new(-det * b, det * a) }) }
    }
    pub fn abs(&self) -> Self { Self { col1: self.col1.abs(), col2: self.col2.abs() } }
}
impl Display for Mat2x2 {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "[{}, {}\n {}, {}]\n", self.col1.x, self.col2.x, self.col1.y, self.col2.y)
    }
}
impl Add for Mat2x2 { type Output = Mat2x2; fn add(self, r: Self) -> Self { Self { col1: self.col1 + r.col1, col2: self.col2 + r.col2 } } }
impl Mul<Vec2> for Mat2x2 {
    type Output = Vec2;
    fn mul(self, r: Vec2) -> Vec2 {
        Vec2 { x: self.col1.x * r.x + self.col2.x * r.y, y: self.col1.y * r.x + self.col2.y * r.y }
    }
}
impl Mul for Mat2x2 { type Output = Mat2x2; fn mul(self, r: Self) -> Self { Self { col1: self * r.col1, col2: self * r.col2 } } }
