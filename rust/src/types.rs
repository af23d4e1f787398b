//! Value types of the wrapper crate's public API: Vec2, Mat2x2, Cross, MathErrors, Sylt2DErrors.
//! Host-side only (the step arithmetic runs on the GPU).  Names, fields and semantics follow the crate being replaced
//! (sylt-2d: math_utils / errors modules) so that callers compile unchanged; lib.rs re-exports them under those paths.
//! `Mat2x2::invert` deliberately keeps that crate's behaviour (adjugate scaled by det), since callers may depend on it.
use crate::arbiter::ArbiterErrors;
use std::fmt::{self, Display, Formatter};
use std::ops::{Add, Mul, Neg, Sub};

#[derive(Debug, Default, PartialOrd, Clone, Copy)]
pub struct Vec2 { pub x: f32, pub y: f32 }

#[derive(Debug, Default, Clone, Copy, PartialEq)]
pub struct Mat2x2 { pub col1: Vec2, pub col2: Vec2 }

#[derive(Debug)]
pub enum MathErrors { NoInverse { matrix: Mat2x2 } }

#[derive(Debug)]
pub enum Sylt2DErrors { MathOperations(MathErrors), Arbiter(ArbiterErrors) }

pub trait Cross<T> { type Output; fn cross(&self, rhs: T) -> Self::Output; }

macro_rules! vec2_binop {
    ($tr:ident, $f:ident, $op:tt) => {
        impl $tr for Vec2 { type Output = Vec2; fn $f(self, r: Vec2) -> Vec2 { Vec2 { x: self.x $op r.x, y: self.y $op r.y } } }
    };
}
vec2_binop!(Add, add, +);
vec2_binop!(Sub, sub, -);
impl Mul<f32> for Vec2 { type Output = Vec2; fn mul(self, k: f32) -> Vec2 { Vec2 { x: self.x * k, y: self.y * k } } }
impl Neg for Vec2 { type Output = Vec2; fn neg(self) -> Vec2 { Vec2 { x: -self.x, y: -self.y } } }
impl PartialEq for Vec2 {
    fn eq(&self, o: &Vec2) -> bool { (self.x - o.x).abs() < f32::EPSILON && (self.y - o.y).abs() < f32::EPSILON }
}
impl From<Vec2> for (f32, f32) { fn from(v: Vec2) -> (f32, f32) { (v.x, v.y) } }
impl Display for Vec2 { fn fmt(&self, f: &mut Formatter<'_>) -> fmt::Result { write!(f, "({}, {})", self.x, self.y) } }

impl Vec2 {
    pub fn new(x: f32, y: f32) -> Vec2 { Vec2 { x, y } }
    pub fn dot(&self, r: Vec2) -> f32 { self.x * r.x + self.y * r.y }
    pub fn length(self) -> f32 { self.dot(self).sqrt() }
    pub fn abs(&self) -> Vec2 { Vec2::new(self.x.abs(), self.y.abs()) }
}
impl Cross<Vec2> for Vec2 { type Output = f32; fn cross(&self, r: Vec2) -> f32 { self.x * r.y - self.y * r.x } }
impl Cross<f32> for Vec2 { type Output = Vec2; fn cross(&self, s: f32) -> Vec2 { Vec2::new(self.y * s, self.x * -s) } }
impl Cross<Vec2> for f32 { type Output = Vec2; fn cross(&self, r: Vec2) -> Vec2 { Vec2::new(-self * r.y, self * r.x) } }

impl Mat2x2 {
    pub fn new(col1: Vec2, col2: Vec2) -> Mat2x2 { Mat2x2 { col1, col2 } }
    pub fn new_from_angle(angle: f32) -> Mat2x2 {
        let (s, c) = (angle.sin(), angle.cos());
        Mat2x2::new(Vec2::new(c, s), Vec2::new(-s, c))
    }
    pub fn transpose(&self) -> Mat2x2 { Mat2x2::new(Vec2::new(self.col1.x, self.col2.x), Vec2::new(self.col1.y, self.col2.y)) }
    pub fn abs(&self) -> Mat2x2 { Mat2x2::new(self.col1.abs(), self.col2.abs()) }
    pub fn invert(&self) -> Result<Mat2x2, MathErrors> {
        let det = self.col1.x * self.col2.y - self.col2.x * self.col1.y;
        if det == 0.0 { return Err(MathErrors::NoInverse { matrix: *self }); }
        Ok(Mat2x2::new(Vec2::new(det * self.col2.y, -det * self.col1.y), Vec2::new(-det * self.col2.x, det * self.col1.x)))
    }
}
impl Add for Mat2x2 { type Output = Mat2x2; fn add(self, r: Mat2x2) -> Mat2x2 { Mat2x2::new(self.col1 + r.col1, self.col2 + r.col2) } }
impl Mul<Vec2> for Mat2x2 {
    type Output = Vec2;
    fn mul(self, v: Vec2) -> Vec2 { Vec2::new(self.col1.x * v.x + self.col2.x * v.y, self.col1.y * v.x + self.col2.y * v.y) }
}
impl Mul for Mat2x2 { type Output = Mat2x2; fn mul(self, r: Mat2x2) -> Mat2x2 { Mat2x2::new(self * r.col1, self * r.col2) } }
impl Display for Mat2x2 {
    fn fmt(&self, f: &mut Formatter<'_>) -> fmt::Result {
        write!(f, "[{}, {}\n {}, {}]\n", self.col1.x, self.col2.x, self.col1.y, self.col2.y)
    }
}

impl Display for MathErrors {
    fn fmt(&self, f: &mut Formatter<'_>) -> fmt::Result {
        let MathErrors::NoInverse { matrix } = self;
        write!(f, "No inverse could be found for matrix {}", matrix)
    }
}
impl Display for Sylt2DErrors {
    fn fmt(&self, f: &mut Formatter<'_>) -> fmt::Result {
        match self {
            Sylt2DErrors::MathOperations(e) => write!(f, "In performing math operations the following error occured: {}", e),
            Sylt2DErrors::Arbiter(e) => write!(f, "In updating and finding the contacts between objects the following error occured: {}", e),
        }
    }
}
impl std::error::Error for MathErrors {}
impl std::error::Error for Sylt2DErrors {}
impl From<MathErrors> for Sylt2DErrors { fn from(e: MathErrors) -> Self { Sylt2DErrors::MathOperations(e) } }
impl From<ArbiterErrors> for Sylt2DErrors { fn from(e: ArbiterErrors) -> Self { Sylt2DErrors::Arbiter(e) } }
