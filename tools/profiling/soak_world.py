"""Long runs of the single-world path (default solver selection): no crashes, no internal errors, finite state.
(A SYLT_ERR_CAPACITY with "why 0" on polygon piles is the documented manifold limit of 16 contacts, not a solver failure:
the reference's polygon contacts carry no position correction, quirk Q-10, so deep piles interpenetrate and explode.)"""
import os, sys, time; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from sylt2d_b200 import scenes, World
cases = (("box_drop 100k", scenes.box_drop(100000, seed=42), 3000), ("box_drop 30k", scenes.box_drop(30000, seed=7), 4000),
         ("mixed_pile 3k", scenes.mixed_pile(3000, seed=9), 2000), ("pyramid 40 rows", scenes.pyramid(40, 10), 3000))
for name, sc, steps in cases:
    w = World(sc.gravity, sc.iterations); w.load_scene(sc)
    t0 = time.time(); done = 0
    try:
        while done < steps:
            w.step_n(sc.dt, 250); done += 250
    except Exception as e:
        print(f"{name}: stopped after {done} steps: {e}", flush=True)
        continue
    st = w.get_bodies(); s = w.get_stats()
    arb, con = w.get_arbiters()
    fin = bool(np.isfinite(st["x"]).all() and np.isfinite(st["y"]).all() and np.isfinite(st["vx"]).all())
    print(f"{name}: {steps} steps in {time.time() - t0:.1f} s, finite={fin}, arbiters={s['arbiters']}, colours={s['colors']}, "
          f"min separation={con['separation'].min():.3f}, |v| 99%={np.percentile(np.hypot(st['vx'], st['vy']), 99):.3f}", flush=True)
