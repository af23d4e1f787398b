//! Sylt2DErrors, as in the reference (src/errors.rs:5-36).
use crate::arbiter::ArbiterErrors;
use crate::math_utils::MathErrors;
use std::fmt;

#[derive(Debug)]
pub enum Sylt2DErrors {
    MathOperations(MathErrors),
    Arbiter(ArbiterErrors),
}
impl fmt::Display for Sylt2DErrors {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            Sylt2DErrors::MathOperations(e) => write!(f, "In performing math operations the following error occured: {}", e),
            Sylt2DErrors::Arbiter(e) => write!(f, "In updating and finding the contacts between objects the following error occured: {}", e),
        }
    }
}
impl std::error::Error for Sylt2DErrors {}
impl From<MathErrors> for Sylt2DErrors { fn from(e: MathErrors) -> Self { Sylt2DErrors::MathOperations(e) } }
impl From<ArbiterErrors> for Sylt2DErrors { fn from(e: ArbiterErrors) -> Self { Sylt2DErrors::Arbiter(e) } }
