"""ctypes binding of the CPU oracle (oracle/liborc.so) — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
The oracle restates /root/reference/src/{world,arbiter,collide,collide_polygon,joint,body,math_utils}.rs;
see oracle/sylt_oracle.cpp for the per-function file:line citations.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BODY_STATE = np.dtype([(n, "<f4") for n in ("x", "y", "rotation", "vx", "vy", "angular_velocity", "fx", "fy", "torque")])
CONTACT = np.dtype([(n, "<f4") for n in ("px", "py", "nx", "ny", "r1x", "r1y", "r2x", "r2y", "separation", "pn", "pt", "pnb",
                                          "mass_normal", "mass_tangent", "bias")]
                   + [(n, "u1") for n in ("in_edge_1", "out_edge_1", "in_edge_2", "out_edge_2")] + [("feature_value", "<i4")])
ARBITER = np.dtype([("id1", "<u8"), ("id2", "<u8"), ("body1", "<i4"), ("body2", "<i4"), ("friction", "<f4"),
                    ("num_contacts", "<i4"), ("contact_offset", "<i4"), ("color", "<i4")])
JOINT_STATE = np.dtype([(n, "<f4") for n in ("px", "py", "bias_x", "bias_y", "r1x", "r1y", "r2x", "r2y", "m11", "m12", "m21", "m22",
                                              "bias_factor", "softness", "la1x", "la1y", "la2x", "la2y")]
                       + [("body1", "<i4"), ("body2", "<i4")])
assert BODY_STATE.itemsize == 36 and CONTACT.itemsize == 68 and ARBITER.itemsize == 40 and JOINT_STATE.itemsize == 80


def build(force=False):
    """Compile liborc.so (and sincos_check) with oracle/Makefile."""
    so = os.path.join(_HERE, "liborc.so")
    src = [os.path.join(_HERE, f) for f in ("sylt_oracle.cpp", "sylt_oracle.h")]
    src.append(os.path.join(_HERE, "..", "sylt2d_b200", "csrc", "sylt_sincos.h"))
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    f32, u64, i32, vp = C.c_float, C.c_uint64, C.c_int32, C.c_void_p
    L.orc_world_create.restype = vp
    L.orc_world_create.argtypes = [f32, f32, C.c_uint32]
    L.orc_world_destroy.argtypes = [vp]
    L.orc_world_clear.argtypes = [vp]
    L.orc_world_set_context.argtypes = [vp, i32, i32, i32]
    L.orc_world_set_mode.argtypes = [vp, i32, i32]
    L.orc_world_set_iterations.argtypes = [vp, C.c_uint32]
    L.orc_world_set_fixes.argtypes = [vp, i32, i32]
    L.orc_world_add_box.argtypes = [vp, u64] + [f32] * 10
    L.orc_world_add_polygon.argtypes = [vp, u64, vp, i32] + [f32] * 8
    L.orc_world_add_joint.argtypes = [vp, u64, u64, f32, f32]
    L.orc_world_set_joint_params.argtypes = [vp, i32, f32, f32]
    L.orc_world_add_force.argtypes = [vp, i32, f32, f32]
    for n in ("num_bodies", "num_joints", "num_arbiters", "num_contacts", "broad_phase"):
        getattr(L, "orc_world_" + n).argtypes = [vp]
    L.orc_world_get_bodies.argtypes = [vp, vp, i32]
    L.orc_world_set_bodies.argtypes = [vp, i32, i32, vp]
    L.orc_world_get_body_props.argtypes = [vp, vp, i32]
    L.orc_world_set_body_props.argtypes = [vp, i32, i32, vp]
    L.orc_world_set_joint_local_anchors.argtypes = [vp, i32, f32, f32, f32, f32]
    L.orc_world_get_arbiters.argtypes = [vp, vp, i32, vp, i32]
    L.orc_world_get_joints.argtypes = [vp, vp, i32]
    L.orc_world_step.argtypes = [vp, f32]
    L.orc_world_step_ordered.argtypes = [vp, f32, vp, i32, vp, i32]
    L.orc_world_last_error_matrix.argtypes = [vp, vp]
    L.orc_collide_pair.argtypes = [i32, i32, f32, f32, vp, i32, f32, f32, f32, i32, f32, f32, vp, i32, f32, f32, f32, vp, i32]
    L.orc_mat_invert.argtypes = [vp, vp]
    L.orc_sincos.argtypes = [i32, f32, vp, vp]
    L.orc_worlds_step_mt.argtypes = [vp, i32, f32, i32, i32]
    _LIB = L
    return L


class OracleError(RuntimeError):
    def __init__(self, code, what=""):
        super().__init__(f"oracle error {code} {what}")
        self.code = code


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleWorld:
    """Mirror of the reference's World (src/world.rs:17-153) on the CPU oracle."""

    def __init__(self, gravity=(0.0, -10.0), iterations=10, sincos_mode=1, broadphase=0):
        self.L = lib()
        self.h = self.L.orc_world_create(gravity[0], gravity[1], iterations)
        self.L.orc_world_set_mode(self.h, sincos_mode, broadphase)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_world_destroy(self.h)
            self.h = None

    def set_context(self, accumulate_impulse=True, warm_starting=False, position_correction=True):
        self.L.orc_world_set_context(self.h, int(accumulate_impulse), int(warm_starting), int(position_correction))

    def set_mode(self, sincos_mode, broadphase):
        self.L.orc_world_set_mode(self.h, sincos_mode, broadphase)

    def set_fixes(self, fix_features=False, fix_inverse=False):
        self.L.orc_world_set_fixes(self.h, int(fix_features), int(fix_inverse))

    def clear(self):
        self.L.orc_world_clear(self.h)

    def add_box(self, id, width, mass, friction=0.0, pos=(0, 0), rot=0.0, vel=(0, 0), omega=0.0):
        r = self.L.orc_world_add_box(self.h, id, width[0], width[1], mass, friction, pos[0], pos[1], rot, vel[0], vel[1], omega)
        if r < 0:
            raise OracleError(-r)
        return r

    def add_polygon(self, id, verts, mass, friction=0.0, pos=(0, 0), rot=0.0, vel=(0, 0), omega=0.0):
        v = np.ascontiguousarray(verts, dtype=np.float32)
        r = self.L.orc_world_add_polygon(self.h, id, _ptr(v), v.shape[0], mass, friction, pos[0], pos[1], rot, vel[0], vel[1], omega)
        if r < 0:
            raise OracleError(-r)
        return r

    def add_joint(self, id1, id2, anchor, softness=None, bias_factor=None):
        j = self.L.orc_world_add_joint(self.h, id1, id2, anchor[0], anchor[1])
        if j < 0:
            raise OracleError(-j)
        if softness is not None or bias_factor is not None:
            self.L.orc_world_set_joint_params(self.h, j, 0.0 if softness is None else softness,
                                              0.2 if bias_factor is None else bias_factor)
        return j

    def set_joint_params(self, joint_index, softness, bias_factor):
        self.L.orc_world_set_joint_params(self.h, int(joint_index), softness, bias_factor)

    def add_force(self, body, f):
        self.L.orc_world_add_force(self.h, body, f[0], f[1])

    def load_scene(self, scene):
        """scene: sylt2d_b200.scenes.Scene (numpy structured arrays)."""
        self.L.orc_world_set_iterations(self.h, scene.iterations)
        self.set_context(*scene.ctx)
        for b in scene.bodies:
            if b["shape"] == 0:
                self.add_box(int(b["id"]), (b["width_x"], b["width_y"]), b["mass"], b["friction"], (b["x"], b["y"]), b["rotation"],
                             (b["vx"], b["vy"]), b["angular_velocity"])
            else:
                v = scene.verts[b["vert_offset"]: b["vert_offset"] + b["vert_count"]]
                self.add_polygon(int(b["id"]), v, b["mass"], b["friction"], (b["x"], b["y"]), b["rotation"],
                                 (b["vx"], b["vy"]), b["angular_velocity"])
        for j in scene.joints:
            self.add_joint(int(j["id1"]), int(j["id2"]), (j["anchor_x"], j["anchor_y"]), j["softness"], j["bias_factor"])

    @property
    def num_bodies(self):
        return self.L.orc_world_num_bodies(self.h)

    @property
    def num_arbiters(self):
        return self.L.orc_world_num_arbiters(self.h)

    @property
    def num_contacts(self):
        return self.L.orc_world_num_contacts(self.h)

    def broad_phase(self):
        e = self.L.orc_world_broad_phase(self.h)
        if e:
            raise OracleError(e)

    def step(self, dt):
        e = self.L.orc_world_step(self.h, dt)
        if e:
            raise OracleError(e)

    def step_ordered(self, dt, keys, joint_order=None):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, 2)
        jo = None if joint_order is None else np.ascontiguousarray(joint_order, dtype=np.int32)
        e = self.L.orc_world_step_ordered(self.h, dt, _ptr(keys), keys.shape[0], _ptr(jo), 0 if jo is None else jo.shape[0])
        if e:
            raise OracleError(e)

    def get_bodies(self):
        out = np.zeros(self.num_bodies, dtype=BODY_STATE)
        self.L.orc_world_get_bodies(self.h, _ptr(out), out.shape[0])
        return out

    def set_bodies(self, states, first=0):
        s = np.ascontiguousarray(states, dtype=BODY_STATE)
        e = self.L.orc_world_set_bodies(self.h, first, s.shape[0], _ptr(s))
        if e:
            raise OracleError(e)

    def get_body_props(self):
        out = np.zeros((self.num_bodies, 8), dtype=np.float32)
        self.L.orc_world_get_body_props(self.h, _ptr(out), out.shape[0])
        return out

    def set_body_props(self, props, first=0):
        p = np.ascontiguousarray(props, dtype=np.float32).reshape(-1, 8)
        e = self.L.orc_world_set_body_props(self.h, first, p.shape[0], _ptr(p))
        if e:
            raise OracleError(e)

    def set_joint_local_anchors(self, joint_index, la1, la2):
        self.L.orc_world_set_joint_local_anchors(self.h, int(joint_index), la1[0], la1[1], la2[0], la2[1])

    def get_arbiters(self):
        na, nc = self.num_arbiters, self.num_contacts
        arb = np.zeros(na, dtype=ARBITER)
        con = np.zeros(nc, dtype=CONTACT)
        r = self.L.orc_world_get_arbiters(self.h, _ptr(arb), na, _ptr(con), nc)
        if r < 0:
            raise OracleError(-r)
        return arb, con

    def get_joints(self):
        out = np.zeros(self.L.orc_world_num_joints(self.h), dtype=JOINT_STATE)
        self.L.orc_world_get_joints(self.h, _ptr(out), out.shape[0])
        return out

    def last_error_matrix(self):
        m = np.zeros(4, dtype=np.float32)
        j = self.L.orc_world_last_error_matrix(self.h, _ptr(m))
        return j, m


def collide_pair(a, b, sincos_mode=1, cap=64):
    """a, b: dicts(shape=0|1, width=(wx,wy), verts=ndarray|None, pos=(x,y), rot=float)."""
    L = lib()
    out = np.zeros(cap, dtype=CONTACT)

    def unpack(d):
        v = None if d.get("verts") is None else np.ascontiguousarray(d["verts"], dtype=np.float32)
        w = d.get("width", (0.0, 0.0))
        return [int(d["shape"]), w[0], w[1], _ptr(v), 0 if v is None else v.shape[0], d["pos"][0], d["pos"][1], d["rot"]], v

    pa, va = unpack(a)
    pb, vb = unpack(b)
    n = L.orc_collide_pair(sincos_mode, *pa, *pb, _ptr(out), cap)
    if n < 0:
        raise OracleError(-n)
    return out[:n].copy()


def mat_invert(m4):
    m = np.ascontiguousarray(m4, dtype=np.float32)
    o = np.zeros(4, dtype=np.float32)
    e = lib().orc_mat_invert(_ptr(m), _ptr(o))
    return e, o


def step_worlds_mt(worlds, dt, steps, threads):
    arr = (C.c_void_p * len(worlds))(*[w.h for w in worlds])
    e = lib().orc_worlds_step_mt(arr, len(worlds), dt, steps, threads)
    if e:
        raise OracleError(e)
