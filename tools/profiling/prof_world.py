import sys, time; sys.path.insert(0,'/root/repo')
from sylt2d_b200 import scenes, World
n = int(sys.argv[1]) if len(sys.argv)>1 else 100000
settle = int(sys.argv[2]) if len(sys.argv)>2 else 600
steps = int(sys.argv[3]) if len(sys.argv)>3 else 3
sc = scenes.box_drop(n, seed=42, iterations=10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
w.step_n(sc.dt, settle)
w.timer_start(); w.step_n_async(sc.dt, steps); ms=w.timer_stop(); w.sync()
print("ms/step", ms/steps, w.get_stats())
