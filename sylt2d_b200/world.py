"""Host-side mirror of the reference's public API for the step path, on top of the C ABI.

Same names, argument meaning and error behaviour as the Rust crate (the Rust toolchain is absent in this image,
so the tests drive the C ABI through this mirror; the Rust wrapper crate itself is under rust/):

    Vec2                      src/math_utils.rs:23-27
    Body.new / new_polygon    src/body.rs:204-284      (ids from a process-global counter starting at 1, :201)
    Joint.new                 src/joint.rs:26-55
    WorldContext              src/world.rs:11-16
    World.new / add_body / add_joint / clear / iter_bodies / broad_phase / step    src/world.rs:36-153
    World.bodies / joints / arbiters  (pub fields, read by examples/samples/src/main.rs:520-588)
"""
import ctypes as C
import itertools
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from ._ffi import ARBITER, BODY_STATE, CONTACT, JOINT_STATE, STEP_STATS, SyltError, check, lib, ptr
from .scenes import BODY_DESC, F32_MAX, JOINT_DESC

_BODY_ID_COUNTER = itertools.count(1)  # BODY_ID_COUNTER (src/body.rs:201)


@dataclass
class Vec2:
    x: float = 0.0
    y: float = 0.0

    def __iter__(self):
        return iter((self.x, self.y))


def _v(t):
    return t if isinstance(t, Vec2) else Vec2(float(t[0]), float(t[1]))


@dataclass
class Body:
    """Construction-time view of a body (src/body.rs:182-199).  After add_body the live state lives on the GPU;
    World.bodies returns fresh snapshots."""
    id: int = 0
    position: Vec2 = field(default_factory=Vec2)
    rotation: float = 0.0
    velocity: Vec2 = field(default_factory=Vec2)
    angular_velocity: float = 0.0
    force: Vec2 = field(default_factory=Vec2)
    torque: float = 0.0
    width: Vec2 = field(default_factory=Vec2)
    friction: float = 0.0
    mass: float = 0.0
    inv_mass: float = 0.0
    moi: float = 0.0
    inv_moi: float = 0.0
    vertices: np.ndarray = None
    shape: int = 0  # 0 Box, 1 ConvexPolygon

    @staticmethod
    def new(width, mass):
        w = _v(width)
        return Body(id=next(_BODY_ID_COUNTER), width=w, mass=float(mass), shape=0)

    @staticmethod
    def new_polygon(vertices, mass):
        v = np.asarray([tuple(p) for p in vertices], dtype=np.float32).reshape(-1, 2)
        return Body(id=next(_BODY_ID_COUNTER), vertices=v, mass=float(mass), shape=1)

    def add_force(self, force):  # src/body.rs:286-288 (before add_body; afterwards use World.add_force)
        f = _v(force)
        self.force = Vec2(float(np.float32(self.force.x) + np.float32(f.x)), float(np.float32(self.force.y) + np.float32(f.y)))


@dataclass
class Joint:
    """Joint::new(body_1, body_2, anchor, &world) (src/joint.rs:26-55); softness / bias_factor are pub fields."""
    body_1: int
    body_2: int
    anchor: Vec2
    softness: float = 0.0
    bias_factor: float = 0.2

    @staticmethod
    def new(body_1, body_2, anchor, world):
        ids = world._ids
        if body_1.id not in ids or body_2.id not in ids:
            raise SyltError(_ffi.ERR_BODY_NOT_FOUND, "couldn't find body in world bodies")  # joint.rs:31,36 `expect`
        return Joint(body_1.id, body_2.id, _v(anchor))


@dataclass
class WorldContext:
    accumulate_impulse: bool = True
    warm_starting: bool = False
    position_correction: bool = True


@dataclass
class Arbiter:
    friction: float
    num_contacts: int
    contacts: np.ndarray  # CONTACT records
    color: int = 0


class World:
    def __init__(self, gravity=(0.0, -10.0), iterations=10, device=0):
        self.L = lib()
        g = _v(gravity)
        h = C.c_void_p()
        check(self.L.sylt_world_create(g.x, g.y, int(iterations), int(device), C.byref(h)))
        self.h = h
        self.world_context = WorldContext()
        self._ids = set()
        self.iterations = int(iterations)

    new = classmethod(lambda cls, gravity, iterations, device=0: cls(gravity, iterations, device))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sylt_world_destroy(self.h)
            self.h = None

    # ------------------------------------------------------------ construction
    def add_body(self, body):
        if body.shape == 0:
            idx = C.c_int32()
            check(self.L.sylt_world_add_box(self.h, body.id, body.width.x, body.width.y, body.mass, body.friction, body.position.x,
                                            body.position.y, body.rotation, body.velocity.x, body.velocity.y, body.angular_velocity,
                                            C.byref(idx)))
        else:
            idx = C.c_int32()
            v = np.ascontiguousarray(body.vertices, dtype=np.float32)
            check(self.L.sylt_world_add_polygon(self.h, body.id, ptr(v), v.shape[0], body.mass, body.friction, body.position.x,
                                                body.position.y, body.rotation, body.velocity.x, body.velocity.y,
                                                body.angular_velocity, C.byref(idx)))
        if body.force.x or body.force.y:
            check(self.L.sylt_world_add_force(self.h, idx.value, body.force.x, body.force.y))
        self._ids.add(body.id)
        return idx.value

    def add_joint(self, joint):
        idx = C.c_int32()
        check(self.L.sylt_world_add_joint(self.h, joint.body_1, joint.body_2, joint.anchor.x, joint.anchor.y, C.byref(idx)))
        check(self.L.sylt_world_set_joint_params(self.h, idx.value, joint.softness, joint.bias_factor))
        return idx.value

    def set_joint_params(self, joint_index, softness, bias_factor):
        """The pub fields Joint.softness / Joint.bias_factor of a joint already in the world (src/joint.rs:17-18)."""
        check(self.L.sylt_world_set_joint_params(self.h, int(joint_index), softness, bias_factor))

    def load_scene(self, scene):
        """Bulk add (sylt_world_add_bodies / add_joints) of a scenes.Scene."""
        check(self.L.sylt_world_set_iterations(self.h, scene.iterations))
        self.iterations = scene.iterations
        self.world_context = WorldContext(*scene.ctx)
        b = np.ascontiguousarray(scene.bodies, dtype=BODY_DESC)
        v = np.ascontiguousarray(scene.verts, dtype=np.float32)
        check(self.L.sylt_world_add_bodies(self.h, ptr(b), b.shape[0], ptr(v) if v.size else None, v.shape[0]))
        self._ids.update(int(i) for i in b["id"])
        j = np.ascontiguousarray(scene.joints, dtype=JOINT_DESC)
        if j.shape[0]:
            check(self.L.sylt_world_add_joints(self.h, ptr(j), j.shape[0]))

    def clear(self):
        check(self.L.sylt_world_clear(self.h))
        self._ids.clear()

    def add_force(self, body_index, force):
        f = _v(force)
        check(self.L.sylt_world_add_force(self.h, body_index, f.x, f.y))

    def set_capacity(self, max_pairs=0, max_arbiters=0, max_contacts=0):
        check(self.L.sylt_world_set_capacity(self.h, max_pairs, max_arbiters, max_contacts))

    def set_fixes(self, fix_features=False, fix_inverse=False):
        """Opt-in corrected physics (default off = the reference's behaviour): see sylt_world_set_fixes."""
        check(self.L.sylt_world_set_fixes(self.h, int(fix_features), int(fix_inverse)))

    def set_iterations(self, iterations):
        check(self.L.sylt_world_set_iterations(self.h, iterations))
        self.iterations = iterations

    # ------------------------------------------------------------ stepping
    def _push_context(self):
        c = self.world_context
        check(self.L.sylt_world_set_context(self.h, int(c.accumulate_impulse), int(c.warm_starting), int(c.position_correction)))

    def broad_phase(self):
        self._push_context()
        check(self.L.sylt_world_broad_phase(self.h))

    def step(self, dt):
        """World::step(dt) -> Result<(), Sylt2DErrors>: raises SyltError on Err (and on the reference's panics)."""
        self._push_context()
        check(self.L.sylt_world_step(self.h, dt))

    def step_n(self, dt, n):
        self._push_context()
        check(self.L.sylt_world_step_n(self.h, dt, n))

    def step_n_async(self, dt, n):
        self._push_context()
        check(self.L.sylt_world_step_n_async(self.h, dt, n))

    def sync(self):
        check(self.L.sylt_world_sync(self.h))

    def timer_start(self):
        check(self.L.sylt_world_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.L.sylt_world_timer_stop(self.h, C.byref(ms)))
        return ms.value

    KERNEL_NAMES = ("prepare", "cells", "fill_cells", "rows_count", "rows_emit", "narrow", "build", "solve")

    def set_kernel_timing(self, on=True):
        """Diagnostics: CUDA events between the launches of every step (serialises the chain of short kernels)."""
        check(self.L.sylt_world_set_kernel_timing(self.h, 1 if on else 0))

    def kernel_times(self):
        """{kernel: microseconds} of the last step's launches (needs set_kernel_timing)."""
        buf = (C.c_float * 16)()
        n = self.L.sylt_world_get_kernel_times(self.h, buf, 16)
        if n < 0:
            check(-n)
        return {self.KERNEL_NAMES[k] if k < len(self.KERNEL_NAMES) else f"k{k}": float(buf[k]) for k in range(n)}

    @property
    def launch_count(self):
        return self.L.sylt_world_launch_count(self.h)

    # ------------------------------------------------------------ read-back
    @property
    def num_bodies(self):
        return self.L.sylt_world_num_bodies(self.h)

    @property
    def num_arbiters(self):
        return self.L.sylt_world_num_arbiters(self.h)

    @property
    def num_contacts(self):
        return self.L.sylt_world_num_contacts(self.h)

    def get_bodies(self):
        out = np.zeros(self.num_bodies, dtype=BODY_STATE)
        r = self.L.sylt_world_get_bodies(self.h, ptr(out), out.shape[0])
        if r < 0:
            raise SyltError(-r)
        return out

    def set_bodies(self, states, first=0):
        s = np.ascontiguousarray(states, dtype=BODY_STATE)
        check(self.L.sylt_world_set_bodies(self.h, first, s.shape[0], ptr(s)))

    def get_body_props(self):
        out = np.zeros((self.num_bodies, 8), dtype=np.float32)
        self.L.sylt_world_get_body_props(self.h, ptr(out), out.shape[0])
        return out

    def set_body_props(self, props, first=0):
        """Write back the (n, 8) records of get_body_props (mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape):
        the pub Body fields the reference reads on every step."""
        p = np.ascontiguousarray(props, dtype=np.float32).reshape(-1, 8)
        check(self.L.sylt_world_set_body_props(self.h, first, p.shape[0], ptr(p)))

    def set_joint_local_anchors(self, joint_index, la1, la2):
        check(self.L.sylt_world_set_joint_local_anchors(self.h, int(joint_index), la1[0], la1[1], la2[0], la2[1]))

    def get_arbiters(self):
        na, nc = self.num_arbiters, self.num_contacts
        arb = np.zeros(na, dtype=ARBITER)
        con = np.zeros(nc, dtype=CONTACT)
        r = self.L.sylt_world_get_arbiters(self.h, ptr(arb), na, ptr(con), nc)
        if r < 0:
            raise SyltError(-r)
        return arb[:r], con

    def get_joints(self):
        out = np.zeros(self.L.sylt_world_num_joints(self.h), dtype=JOINT_STATE)
        r = self.L.sylt_world_get_joints(self.h, ptr(out), out.shape[0])
        if r < 0:
            raise SyltError(-r)
        return out

    def get_joint_order(self):
        out = np.zeros(self.L.sylt_world_num_joints(self.h), dtype=np.int32)
        r = self.L.sylt_world_get_joint_order(self.h, ptr(out), out.shape[0])
        if r < 0:
            raise SyltError(-r)
        return out

    def get_stats(self):
        out = np.zeros(1, dtype=STEP_STATS)
        check(self.L.sylt_world_get_stats(self.h, ptr(out)))
        return {k: int(out[0][k]) for k in STEP_STATS.names}

    def solve_order(self):
        """(keys, joint_order) of the last step: arbiter (id1,id2) keys in the order the GPU solved them (colour-major)."""
        arb, _ = self.get_arbiters()
        o = np.lexsort((arb["id2"], arb["id1"], arb["color"]))
        keys = np.stack([arb["id1"][o], arb["id2"][o]], axis=1).astype(np.uint64)
        return keys, self.get_joint_order()

    def last_error_matrix(self):
        m = np.zeros(4, dtype=np.float32)
        j = self.L.sylt_world_last_error_matrix(self.h, ptr(m))
        return j, m

    # pub-field mirrors of the reference ------------------------------------------------
    def iter_bodies(self):
        return iter(self.bodies)

    @property
    def bodies(self):
        st, pr = self.get_bodies(), self.get_body_props()
        return [Body(id=0, position=Vec2(float(s["x"]), float(s["y"])), rotation=float(s["rotation"]),
                     velocity=Vec2(float(s["vx"]), float(s["vy"])), angular_velocity=float(s["angular_velocity"]),
                     force=Vec2(float(s["fx"]), float(s["fy"])), torque=float(s["torque"]), width=Vec2(float(p[4]), float(p[5])),
                     friction=float(p[6]), mass=float(p[0]), inv_mass=float(p[1]), moi=float(p[2]), inv_moi=float(p[3]), shape=int(p[7]))
                for s, p in zip(st, pr)]

    @property
    def arbiters(self):
        arb, con = self.get_arbiters()
        return {(int(a["id1"]), int(a["id2"])): Arbiter(float(a["friction"]), int(a["num_contacts"]),
                                                        con[a["contact_offset"]: a["contact_offset"] + a["num_contacts"]], int(a["color"]))
                for a in arb}

    @property
    def joints(self):
        return self.get_joints()


def collide_pairs(a, b, verts=None, cap=16, device=0):
    """Narrow phase on explicit pairs (collide / collide_polygons): a, b are BODY_DESC arrays of equal length."""
    a = np.ascontiguousarray(a, dtype=BODY_DESC)
    b = np.ascontiguousarray(b, dtype=BODY_DESC)
    v = None if verts is None else np.ascontiguousarray(verts, dtype=np.float32)
    n = a.shape[0]
    out = np.zeros((n, cap), dtype=CONTACT)
    cnt = np.zeros(n, dtype=np.int32)
    check(lib().sylt_collide_pairs(device, ptr(a), ptr(b), n, ptr(v), 0 if v is None else v.shape[0], ptr(out), cap, ptr(cnt)))
    return out, cnt


def device_sincos(angles, device=0):
    a = np.ascontiguousarray(angles, dtype=np.float32)
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    check(lib().sylt_sincos(device, ptr(a), a.shape[0], ptr(s), ptr(c)))
    return s, c
