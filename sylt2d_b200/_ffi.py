"""ctypes binding of libsylt2d_b200.so (the C ABI declared in include/sylt2d_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the symbols
(`lib()`), and every compute entry point returns SYLT_ERR_CUDA when no device is usable.
"""
import ctypes as C
import os

import numpy as np

from .scenes import BODY_DESC, JOINT_DESC

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYLT2D_B200_LIB: load another build of the same library (profiling builds under tools/profiling)
LIB_PATH = os.environ.get("SYLT2D_B200_LIB") or os.path.join(_HERE, "libsylt2d_b200.so")

SYLT_OK, ERR_NO_INVERSE, ERR_ARBITER, ERR_BODY_ORDER, ERR_BODY_NOT_FOUND, ERR_CAPACITY, ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED = range(9)
ERROR_NAMES = {1: "NoInverse (Sylt2DErrors::MathOperations)", 2: "Arbiter", 3: "BodyOrder", 4: "BodyNotFound", 5: "Capacity",
               6: "CUDA", 7: "Invalid", 8: "Unsupported", 9: "Internal (solver watchdog)"}

BODY_STATE = np.dtype([(n, "<f4") for n in ("x", "y", "rotation", "vx", "vy", "angular_velocity", "fx", "fy", "torque")])
CONTACT = np.dtype([(n, "<f4") for n in ("px", "py", "nx", "ny", "r1x", "r1y", "r2x", "r2y", "separation", "pn", "pt", "pnb",
                                          "mass_normal", "mass_tangent", "bias")]
                   + [(n, "u1") for n in ("in_edge_1", "out_edge_1", "in_edge_2", "out_edge_2")] + [("feature_value", "<i4")])
ARBITER = np.dtype([("id1", "<u8"), ("id2", "<u8"), ("body1", "<i4"), ("body2", "<i4"), ("friction", "<f4"),
                    ("num_contacts", "<i4"), ("contact_offset", "<i4"), ("color", "<i4")])
JOINT_STATE = np.dtype([(n, "<f4") for n in ("px", "py", "bias_x", "bias_y", "r1x", "r1y", "r2x", "r2y", "m11", "m12", "m21", "m22",
                                              "bias_factor", "softness", "la1x", "la1y", "la2x", "la2y")]
                       + [("body1", "<i4"), ("body2", "<i4")])
STEP_STATS = np.dtype([(n, "<i4") for n in ("bodies", "candidate_pairs", "arbiters", "contacts", "joints", "colors", "joint_colors", "color_rounds")])
BATCH_STATS = np.dtype([("body_steps", "<u8"), ("contacts", "<u8"), ("errors", "<u8"), ("kinetic_energy", "<f8"),
                        ("max_penetration", "<f4"), ("_pad", "<f4")])
assert BODY_STATE.itemsize == 36 and CONTACT.itemsize == 68 and ARBITER.itemsize == 40 and JOINT_STATE.itemsize == 80
assert STEP_STATS.itemsize == 32 and BATCH_STATS.itemsize == 40

# every symbol include/sylt2d_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "sylt_last_error_string", "sylt_device_count",
    "sylt_world_create", "sylt_world_destroy", "sylt_world_clear", "sylt_world_set_context", "sylt_world_set_iterations",
    "sylt_world_set_capacity", "sylt_world_set_fixes", "sylt_world_add_box", "sylt_world_add_polygon", "sylt_world_add_bodies", "sylt_world_add_joint",
    "sylt_world_add_joints", "sylt_world_set_joint_params", "sylt_world_add_joint_local", "sylt_world_set_joint_local_anchors", "sylt_world_set_body_props", "sylt_world_add_force", "sylt_world_num_bodies", "sylt_world_num_joints",
    "sylt_world_num_arbiters", "sylt_world_num_contacts", "sylt_world_get_bodies", "sylt_world_set_bodies", "sylt_world_get_body_props",
    "sylt_world_get_arbiters", "sylt_world_get_joints", "sylt_world_get_joint_order", "sylt_world_get_stats",
    "sylt_world_last_error_matrix", "sylt_world_broad_phase", "sylt_world_step", "sylt_world_step_n", "sylt_world_step_n_async",
    "sylt_world_sync", "sylt_world_timer_start", "sylt_world_timer_stop", "sylt_world_launch_count", "sylt_world_set_kernel_timing", "sylt_world_get_kernel_times", "sylt_collide_pairs", "sylt_sincos",
    "sylt_batch_create", "sylt_batch_destroy", "sylt_batch_set_context", "sylt_batch_set_capacity", "sylt_batch_set_export_worlds", "sylt_batch_set_fixes", "sylt_batch_step_n", "sylt_batch_step_n_async",
    "sylt_batch_sync", "sylt_batch_step_host", "sylt_batch_step_host6", "sylt_batch_timer_start", "sylt_batch_timer_stop", "sylt_batch_launch_count", "sylt_batch_num_worlds",
    "sylt_batch_num_bodies", "sylt_batch_get_bodies", "sylt_batch_set_bodies", "sylt_batch_get_arbiters", "sylt_batch_get_joints",
    "sylt_batch_get_joint_order", "sylt_batch_get_errors", "sylt_batch_stats",
]

_LIB = None


class SyltError(RuntimeError):
    """Mirrors Sylt2DErrors (src/errors.rs:5-9) plus the conditions the reference panics on."""

    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"sylt2d_b200 error {code} ({ERROR_NAMES.get(code, '?')}) {detail}")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    f32, u64, i32, i64, vp, u32 = C.c_float, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p, C.c_uint32
    L.sylt_last_error_string.restype = C.c_char_p
    L.sylt_world_create.argtypes = [f32, f32, u32, i32, vp]
    for n in ("destroy", "clear", "num_bodies", "num_joints", "num_arbiters", "num_contacts", "broad_phase", "sync", "timer_start"):
        getattr(L, "sylt_world_" + n).argtypes = [vp]
    L.sylt_world_set_context.argtypes = [vp, i32, i32, i32]
    L.sylt_world_set_iterations.argtypes = [vp, u32]
    L.sylt_world_set_fixes.argtypes = [vp, i32, i32]
    L.sylt_batch_set_fixes.argtypes = [vp, i32, i32]
    L.sylt_batch_set_capacity.argtypes = [vp, i32]
    L.sylt_batch_set_export_worlds.argtypes = [vp, vp, i32]
    L.sylt_world_set_capacity.argtypes = [vp, i64, i64, i64]
    L.sylt_world_add_box.argtypes = [vp, u64] + [f32] * 10 + [vp]
    L.sylt_world_add_polygon.argtypes = [vp, u64, vp, i32] + [f32] * 8 + [vp]
    L.sylt_world_add_bodies.argtypes = [vp, vp, i32, vp, i32]
    L.sylt_world_add_joint.argtypes = [vp, u64, u64, f32, f32, vp]
    L.sylt_world_add_joints.argtypes = [vp, vp, i32]
    L.sylt_world_set_joint_params.argtypes = [vp, i32, f32, f32]
    L.sylt_world_add_joint_local.argtypes = [vp, u64, u64, f32, f32, f32, f32, f32, f32, vp]
    L.sylt_world_set_joint_local_anchors.argtypes = [vp, i32, f32, f32, f32, f32]
    L.sylt_world_set_body_props.argtypes = [vp, i32, i32, vp]
    L.sylt_world_add_force.argtypes = [vp, i32, f32, f32]
    L.sylt_world_get_bodies.argtypes = [vp, vp, i32]
    L.sylt_world_set_bodies.argtypes = [vp, i32, i32, vp]
    L.sylt_world_get_body_props.argtypes = [vp, vp, i32]
    L.sylt_world_get_arbiters.argtypes = [vp, vp, i32, vp, i32]
    L.sylt_world_get_joints.argtypes = [vp, vp, i32]
    L.sylt_world_get_joint_order.argtypes = [vp, vp, i32]
    L.sylt_world_get_stats.argtypes = [vp, vp]
    L.sylt_world_last_error_matrix.argtypes = [vp, vp]
    L.sylt_world_step.argtypes = [vp, f32]
    L.sylt_world_step_n.argtypes = [vp, f32, i32]
    L.sylt_world_step_n_async.argtypes = [vp, f32, i32]
    L.sylt_world_timer_stop.argtypes = [vp, vp]
    L.sylt_world_set_kernel_timing.argtypes = [vp, C.c_int]
    L.sylt_world_get_kernel_times.argtypes = [vp, vp, C.c_int32]
    L.sylt_world_launch_count.argtypes = [vp]
    L.sylt_world_launch_count.restype = i64
    L.sylt_collide_pairs.argtypes = [i32, vp, vp, i32, vp, i32, vp, i32, vp]
    L.sylt_sincos.argtypes = [i32, vp, i32, vp, vp]
    L.sylt_batch_create.argtypes = [i32, vp, vp, vp, vp, f32, f32, u32, i32, vp]
    for n in ("destroy", "sync", "timer_start", "num_worlds"):
        getattr(L, "sylt_batch_" + n).argtypes = [vp]
    L.sylt_batch_set_context.argtypes = [vp, i32, i32, i32]
    L.sylt_batch_step_n.argtypes = [vp, f32, i32]
    L.sylt_batch_step_n_async.argtypes = [vp, f32, i32]
    L.sylt_batch_timer_stop.argtypes = [vp, vp]
    L.sylt_batch_step_host.argtypes = [vp, f32, i32, vp, vp]
    L.sylt_batch_step_host6.argtypes = [vp, f32, i32, vp, vp]
    L.sylt_batch_launch_count.argtypes = [vp]
    L.sylt_batch_launch_count.restype = i64
    L.sylt_batch_num_bodies.argtypes = [vp]
    L.sylt_batch_num_bodies.restype = i64
    L.sylt_batch_get_bodies.argtypes = [vp, i32, i32, vp, i64]
    L.sylt_batch_set_bodies.argtypes = [vp, i32, i32, vp, i64]
    L.sylt_batch_get_arbiters.argtypes = [vp, i32, vp, i32, vp, i32]
    L.sylt_batch_get_joints.argtypes = [vp, i32, vp, i32]
    L.sylt_batch_get_joint_order.argtypes = [vp, i32, vp, i32]
    L.sylt_batch_get_errors.argtypes = [vp, vp, i32]
    L.sylt_batch_stats.argtypes = [vp, vp]
    _LIB = L
    return L


def check(code):
    if code != SYLT_OK:
        raise SyltError(code, lib().sylt_last_error_string().decode(errors="replace") if code == ERR_CUDA else "")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


__all__ = ["lib", "check", "ptr", "SyltError", "BODY_DESC", "JOINT_DESC", "BODY_STATE", "CONTACT", "ARBITER", "JOINT_STATE",
           "STEP_STATS", "BATCH_STATS", "SYMBOLS", "LIB_PATH"]
