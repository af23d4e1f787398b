"""Per-phase cycle breakdown of k_batch_step (profiling build with -DSYLT_BATCH_TIMING).

  here (no GPU):  python tools/profiling/batch_phases.py --build      # writes scratch/libsylt2d_b200_timing.so
  on the B200:    python tools/profiling/batch_phases.py [worlds] [steps]

Thread 0 of every CTA accumulates clock64() deltas per phase; the numbers are latencies of a CTA that shares its SM with
three others, so they add up to the CTA's residency time, not to SM-exclusive time."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "scratch", "libsylt2d_b200_timing.so")
sys.path.insert(0, ROOT)

if "--build" in sys.argv:
    from sylt2d_b200 import build as B
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["nvcc"] + B.NVCC_FLAGS + ["-DSYLT_BATCH_TIMING=" + ("2" if "--hops" in sys.argv else "1"), "-o", LIB] + B.SOURCES, cwd=B.CSRC, check=True)
    print(LIB)
    sys.exit(0)

os.environ["SYLT2D_B200_LIB"] = LIB
import numpy as np
from sylt2d_b200 import Batch, scenes
from sylt2d_b200._ffi import lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
worlds = int(args[0]) if args else 4096
steps = int(args[1]) if len(args) > 1 else 20
sc = scenes.pyramid(20, 10)
bodies, boff, joints, joff = scenes.tile_worlds(sc, worlds, seed=0, x_jitter=0.01)
b = Batch(bodies, boff, joints[:0], joff, sc.gravity, sc.iterations, device=0, ctx=sc.ctx)
b.step_n(sc.dt, 150)
L = lib()
out = (C.c_ulonglong * 16)()
L.sylt_batch_debug_phases(out, 1)
b.timer_start(); b.step_n_async(sc.dt, steps); ms = b.timer_stop(); b.sync()
L.sylt_batch_debug_phases(out, 0)
names = ["loop head", "1 bodies", "2 broad phase", "3 narrow+match", "4 colouring", "5 arbiter pre_step", "6 joint pre_step", "7 iterations",
         "8 integrate+cache", "2a bucket count", "2b bucket scan+fill", "2c row search", "2d row scan"]
names[2] = "2e row emit+sort"
tot = sum(out[:13])
print(f"{worlds} worlds x {steps} steps: {ms / steps:.3f} ms/step, {worlds * sc.num_bodies * steps / ms * 1e3:.3e} body-steps/s")
for i, n in enumerate(names):
    print(f"  {n:20s} {out[i] / (worlds * steps):10.0f} cycles/world-step  {100.0 * out[i] / tot:5.1f} %")
print(f"  total {tot / (worlds * steps):.0f} cycles/world-step")
hops = worlds * steps * sc.iterations
print("inside the iteration phase, thread 0, cycles per ITERATION (sum over its colours):")
for i, n in ((13, "hop: indices + body loads"), (14, "hop: contact loads + arithmetic"), (15, "hop: stores + barrier")):
    print(f"  {n:34s} {out[i] / hops:8.0f}")
