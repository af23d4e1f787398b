import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from sylt2d_b200 import Batch, scenes
sc = scenes.pyramid(20, 10)
n=32768
bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=0, x_jitter=0.01)
b = Batch(bodies, boff, joints[:0], joff, sc.gravity, sc.iterations, ctx=sc.ctx)
b.step_n(sc.dt, 150)
for per in (1, 2, 5, 10, 20):
    b.step_n(sc.dt, per)
    b.timer_start()
    for _ in range(20//per): b.step_n_async(sc.dt, per)
    ms=b.timer_stop()
    print("steps/launch", per, "ms/step", ms/20, "body-steps/s", n*211*20/(ms*1e-3))
