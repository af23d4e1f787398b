//! sylt-2d with World::step on a B200 (libsylt2d_b200).  Public module paths are those of the crate being replaced
//! (src/lib.rs:1-9 there); `draw` (terminal ASCII renderer) is out of scope.
//! NOT COMPILED in this repository's environment (no rustc): see INTEGRATION.md.
pub mod arbiter;
pub mod body;
pub mod collide;
pub mod collide_polygon;
pub mod ffi;
pub mod joint;
mod types;
pub mod world;
pub mod math_utils { pub use crate::types::{Cross, Mat2x2, MathErrors, Vec2}; }
pub mod errors { pub use crate::types::Sylt2DErrors; }
