"""How long would a cached candidate-pair list stay valid?  (design input for the temporal-coherence broad phase)
A list built from AABBs inflated by m stays a superset of the true candidates while every body's current AABB lies inside
its inflated AABB of build time.  Emulates that rule on the settled 100k-body world and counts the rebuilds.
usage: python tools/profiling/settle_motion.py [bodies] [settle] [steps]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from sylt2d_b200 import World, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 600
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
sc = scenes.box_drop(n, seed=42)
w = World(sc.gravity, sc.iterations)
w.load_scene(sc)
props = None


def aabbs(st):
    c, s = np.abs(np.cos(st["rotation"])), np.abs(np.sin(st["rotation"]))
    hx, hy = props[:, 4] * 0.5, props[:, 5] * 0.5
    ex, ey = c * hx + s * hy, s * hx + c * hy
    return st["x"] - ex, st["y"] - ey, st["x"] + ex, st["y"] + ey


for phase, pre in (("falling", 5), ("settled", settle)):
    w.step_n(sc.dt, pre)
    props = w.get_body_props()
    dyn = props[:, 1] != 0
    margins = [0.02, 0.05, 0.1, 0.2]
    ref = {m: aabbs(w.get_bodies()) for m in margins}
    rebuilds = {m: 0 for m in margins}
    viol = {m: [] for m in margins}
    for s in range(steps):
        w.step(sc.dt)
        a = aabbs(w.get_bodies())
        for m in margins:
            r = ref[m]
            out = (a[0] < r[0] - m) | (a[1] < r[1] - m) | (a[2] > r[2] + m) | (a[3] > r[3] + m)
            out &= dyn
            if out.any():
                rebuilds[m] += 1
                viol[m].append(int(out.sum()))
                ref[m] = a
    st = w.get_bodies()
    sp = np.sqrt(st["vx"] ** 2 + st["vy"] ** 2)[dyn]
    print(phase, "steps", steps, "speed percentiles 50/99/99.9/max", np.percentile(sp, [50, 99, 99.9, 100]))
    for m in margins:
        print(f"  margin {m}: {rebuilds[m]} rebuilds in {steps} steps; bodies out at a rebuild: median {np.median(viol[m]) if viol[m] else 0}, max {max(viol[m]) if viol[m] else 0}")
    print("  stats", w.get_stats())
