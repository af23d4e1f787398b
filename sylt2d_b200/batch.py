"""Batched independent worlds on one GPU (one CTA per world) — wrapper of the sylt_batch_* C ABI.

The reference has no batched API: a batch is N separate `World`s (src/world.rs:17-153) stepped together.
Multi-GPU: one process per GPU, each owning a contiguous shard of the worlds (`shard_range`); nothing is exchanged
per step, only the final SyltBatchStats record is reduced by the caller (bench.py uses torch.distributed for that).
"""
import ctypes as C

import numpy as np

from ._ffi import ARBITER, BATCH_STATS, BODY_STATE, CONTACT, JOINT_STATE, SyltError, check, lib, ptr
from .scenes import BODY_DESC, JOINT_DESC


def shard_range(n_worlds, rank, world_size):
    """Contiguous block partition of worlds over ranks."""
    per = (n_worlds + world_size - 1) // world_size
    lo = min(n_worlds, rank * per)
    return lo, min(n_worlds, lo + per)


class Batch:
    def __init__(self, bodies, body_offsets, joints=None, joint_offsets=None, gravity=(0.0, -10.0), iterations=10, device=0,
                 ctx=None):
        """ctx = (accumulate_impulse, warm_starting, position_correction); None keeps World::new's defaults
        (true, false, true; src/world.rs:38-42), exactly like N default reference Worlds."""
        self.L = lib()
        self.bodies = np.ascontiguousarray(bodies, dtype=BODY_DESC)
        self.body_offsets = np.ascontiguousarray(body_offsets, dtype=np.int32)
        n = self.body_offsets.shape[0] - 1
        if joints is None or joint_offsets is None:
            joints = np.zeros(0, dtype=JOINT_DESC)
            joint_offsets = np.zeros(n + 1, dtype=np.int32)
        self.joints = np.ascontiguousarray(joints, dtype=JOINT_DESC)
        self.joint_offsets = np.ascontiguousarray(joint_offsets, dtype=np.int32)
        h = C.c_void_p()
        check(self.L.sylt_batch_create(n, ptr(self.body_offsets), ptr(self.bodies), ptr(self.joint_offsets),
                                       ptr(self.joints) if self.joints.shape[0] else None, gravity[0], gravity[1], int(iterations),
                                       int(device), C.byref(h)))
        self.h = h
        self.n_worlds = n
        if ctx is not None:
            self.set_context(*ctx)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sylt_batch_destroy(self.h)
            self.h = None

    def set_context(self, accumulate_impulse=True, warm_starting=False, position_correction=True):
        check(self.L.sylt_batch_set_context(self.h, int(accumulate_impulse), int(warm_starting), int(position_correction)))

    def step_host6(self, dt, n, states_in=None, states_out=None):
        """step_host with compact float32 (n_bodies, 6) arrays: x, y, rotation, vx, vy, angular_velocity."""
        nb = self.num_bodies
        for a in (states_in, states_out):
            assert a is None or (a.dtype == np.float32 and a.shape == (nb, 6) and a.flags["C_CONTIGUOUS"])
        check(self.L.sylt_batch_step_host6(self.h, dt, n, ptr(states_in), ptr(states_out)))
        return states_out

    def set_capacity(self, slots_per_world=0):
        """Candidate-pair slots per world (0 = default 2 * max bodies + 32); drops the contact history."""
        check(self.L.sylt_batch_set_capacity(self.h, int(slots_per_world)))

    def set_export_worlds(self, worlds):
        """Worlds (<= 64) whose arbiters get_arbiters / solve_order can read back after the next step."""
        w = np.ascontiguousarray(worlds, dtype=np.int32)
        check(self.L.sylt_batch_set_export_worlds(self.h, ptr(w), w.shape[0]))

    def set_fixes(self, fix_features=False, fix_inverse=False):
        check(self.L.sylt_batch_set_fixes(self.h, int(fix_features), int(fix_inverse)))

    def step_n(self, dt, n):
        check(self.L.sylt_batch_step_n(self.h, dt, n))

    def step_n_async(self, dt, n):
        check(self.L.sylt_batch_step_n_async(self.h, dt, n))

    def sync(self):
        check(self.L.sylt_batch_sync(self.h))

    def step_host(self, dt, n, states_in=None, states_out=None):
        """set_bodies + step_n + get_bodies in one pipelined call on HOST arrays (BODY_STATE, ideally pinned)."""
        nb = self.num_bodies
        for a in (states_in, states_out):
            assert a is None or (a.dtype == BODY_STATE and a.shape[0] == nb and a.flags["C_CONTIGUOUS"])
        check(self.L.sylt_batch_step_host(self.h, dt, n, ptr(states_in), ptr(states_out)))
        return states_out

    def timer_start(self):
        check(self.L.sylt_batch_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.L.sylt_batch_timer_stop(self.h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return self.L.sylt_batch_launch_count(self.h)

    @property
    def num_bodies(self):
        return self.L.sylt_batch_num_bodies(self.h)

    def get_bodies(self, first_world=0, n_worlds=None):
        n_worlds = self.n_worlds - first_world if n_worlds is None else n_worlds
        n = int(self.body_offsets[first_world + n_worlds] - self.body_offsets[first_world])
        out = np.zeros(n, dtype=BODY_STATE)
        r = self.L.sylt_batch_get_bodies(self.h, first_world, n_worlds, ptr(out), n)
        if r < 0:
            raise SyltError(-r)
        return out

    def set_bodies(self, states, first_world=0, n_worlds=None):
        n_worlds = self.n_worlds - first_world if n_worlds is None else n_worlds
        s = np.ascontiguousarray(states, dtype=BODY_STATE)
        check(self.L.sylt_batch_set_bodies(self.h, first_world, n_worlds, ptr(s), s.shape[0]))

    def get_arbiters(self, world, cap=4096):
        arb = np.zeros(cap, dtype=ARBITER)
        con = np.zeros(2 * cap, dtype=CONTACT)
        r = self.L.sylt_batch_get_arbiters(self.h, world, ptr(arb), cap, ptr(con), 2 * cap)
        if r < 0:
            raise SyltError(-r)
        arb = arb[:r]
        nc = int(arb["num_contacts"].sum())
        return arb, con[:nc]

    def get_joints(self, world):
        n = int(self.joint_offsets[world + 1] - self.joint_offsets[world])
        out = np.zeros(n, dtype=JOINT_STATE)
        r = self.L.sylt_batch_get_joints(self.h, world, ptr(out), n)
        if r < 0:
            raise SyltError(-r)
        return out

    def get_joint_order(self, world):
        n = int(self.joint_offsets[world + 1] - self.joint_offsets[world])
        out = np.zeros(n, dtype=np.int32)
        r = self.L.sylt_batch_get_joint_order(self.h, world, ptr(out), n)
        if r < 0:
            raise SyltError(-r)
        return out

    def solve_order(self, world):
        arb, _ = self.get_arbiters(world)
        o = np.lexsort((arb["id2"], arb["id1"], arb["color"]))
        keys = np.stack([arb["id1"][o], arb["id2"][o]], axis=1).astype(np.uint64)
        return keys, self.get_joint_order(world)

    def get_errors(self):
        out = np.zeros(self.n_worlds, dtype=np.int32)
        r = self.L.sylt_batch_get_errors(self.h, ptr(out), out.shape[0])
        if r < 0:
            raise SyltError(-r)
        return out

    def stats(self):
        out = np.zeros(1, dtype=BATCH_STATS)
        check(self.L.sylt_batch_stats(self.h, ptr(out)))
        return {k: out[0][k].item() for k in BATCH_STATS.names if k != "_pad"}
