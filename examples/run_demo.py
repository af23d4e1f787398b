#!/usr/bin/env python
"""Headless runner of the reference's sample scenes (examples/samples/src/main.rs demo1..demo10) on the B200 backend:
load a demo, step it, optionally launch "bombs" between steps (launch_bomb, main.rs:56-64), print a summary.
usage: python examples/run_demo.py demo5 --steps 600 [--bombs 3] [--iterations 100]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sylt2d_b200 import Body, Vec2, World, scenes  # noqa: E402

DEMOS = {"demo1": scenes.demo1, "demo2": scenes.demo2, "demo3": scenes.demo3, "demo4": scenes.demo4, "demo5": lambda it: scenes.pyramid(12, it),
         "demo6": scenes.demo6, "demo7": scenes.bridge, "demo8": scenes.demo8, "demo9": scenes.multi_pendulum, "demo10": scenes.demo10}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("demo", choices=sorted(DEMOS))
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--iterations", type=int, default=100)  # ITERATIONS (main.rs:10)
    ap.add_argument("--bombs", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    sc = DEMOS[a.demo](a.iterations)
    w = World(sc.gravity, sc.iterations)
    w.step(sc.dt)  # the sample steps the empty world once before loading a demo (main.rs:394-401)
    w.load_scene(sc)
    rng = np.random.default_rng(a.seed)
    bomb_at = set(np.linspace(a.steps // 4, a.steps - 1, a.bombs, dtype=int).tolist()) if a.bombs else set()
    nid = int(sc.bodies["id"].max()) + 1
    for s in range(a.steps):
        if s in bomb_at:
            b = Body.new(Vec2(1.0, 1.0), 50.0)
            b.id, nid = nid, nid + 1
            b.friction = 0.2
            b.position = Vec2(float(rng.uniform(-15, 15)), 15.0)
            b.rotation = float(rng.uniform(-1.5, 1.5))
            b.velocity = Vec2(b.position.x * -1.5, b.position.y * -1.5)
            b.angular_velocity = float(rng.uniform(-20, 20))
            w.add_body(b)
        w.step(sc.dt)
    st, stats = w.get_bodies(), w.get_stats()
    print(f"{a.demo}: bodies {stats['bodies']} arbiters {stats['arbiters']} contacts {stats['contacts']} joints {stats['joints']} "
          f"colours {stats['colors']}+{stats['joint_colors']}  y range [{st['y'].min():.3f}, {st['y'].max():.3f}]  "
          f"max speed {np.sqrt(st['vx'] ** 2 + st['vy'] ** 2).max():.3f}")


if __name__ == "__main__":
    main()
