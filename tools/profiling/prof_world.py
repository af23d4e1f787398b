"""ms/step of the box-drop world after `settle` steps; SYLT_SOLVER_TIMING=1 prints k_solve's section times to stderr."""
import os, sys; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from sylt2d_b200 import scenes, World
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 600
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sc = scenes.box_drop(n, seed=42, iterations=10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
done = 0
while done < settle:  # several calls: the region partition is only refreshed between step_n calls
    k = min(100, settle - done); w.step_n(sc.dt, k); done += k
for r in range(reps):
    w.timer_start(); w.step_n_async(sc.dt, steps); ms = w.timer_stop(); w.sync()
    print("ms/step", ms / steps, w.get_stats(), flush=True)
w.step_n(sc.dt, 1)
