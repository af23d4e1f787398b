//! sylt-2d with World::step on a B200 (libsylt2d_b200).  Module layout mirrors the reference crate
//! (/root/reference/src/lib.rs:1-9); `draw` (terminal ASCII renderer) is out of scope.
pub mod arbiter;
pub mod body;
pub mod errors;
pub mod ffi;
pub mod joint;
pub mod math_utils;
pub mod world;
