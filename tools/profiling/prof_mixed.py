"""ms/step of the 10k mixed box + polygon pile (BASELINE.json config 2) after `settle` steps."""
import os, sys; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from sylt2d_b200 import scenes, World
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 150
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenes.mixed_pile(10000, seed=1234, iterations=10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
w.step_n(sc.dt, settle)
w.timer_start(); w.step_n_async(sc.dt, steps); ms = w.timer_stop(); w.sync()
print("ms/step", ms / steps, w.get_stats(), flush=True)
