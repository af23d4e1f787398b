import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from sylt2d_b200 import scenes, World
sc = scenes.box_drop(100000, seed=42, iterations=10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
w.step_n(sc.dt, 640)
arb,_ = w.get_arbiters()
c = arb["color"]
print("A", len(arb), "boundary frac", (c>=40).mean())
print("interior colour hist", np.bincount(c[c<40], minlength=12)[:16])
print("boundary colour hist", np.bincount(c[c>=40]-40, minlength=12)[:16])
