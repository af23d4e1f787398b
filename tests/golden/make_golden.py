"""Generates the golden fixtures in this directory (run from the repo root: `python tests/golden/make_golden.py`).

The reference is a Rust crate and there is no Rust toolchain in the build image, so these vectors do NOT come from the
reference binary (parity stays "unpinned by the reference").  They are outputs of the C++ oracle
(oracle/sylt_oracle.cpp, glibc-faithful trig) that the independent numpy-f32 restatement (oracle/narrow_np.py,
oracle/step_np.py) reproduces bit for bit at generation time — both restatements were written separately from the Rust
sources.  Committed so that (a) a later change of either restatement or of the shared sincos cannot move the answers
unnoticed, (b) the CUDA path is checked against fixed bytes, not only against an oracle built on the same box.

  narrow_pairs.npz   600 seeded box / polygon pairs (collide.rs:140-297, collide_polygon.rs:195-203): inputs + contacts
  steps.npz          body states + arbiter keys + accumulated impulses after N World::step calls in canonical
                     (id1, id2) arbiter order: pyramid(5), a mixed box / polygon scene, the bridge, the multi-pendulum,
                     the single pendulum (demo2)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
CONTACT_FIELDS = ("px", "py", "nx", "ny", "separation", "in_edge_1", "out_edge_1", "in_edge_2", "out_edge_2")
STATE_FIELDS = ("x", "y", "rotation", "vx", "vy", "angular_velocity")


def narrow_inputs(n=600, seed=11):
    """(kind, params) per pair; kind 0 = box/box, 1 = polygon/polygon, 2 = box/polygon."""
    rng = np.random.default_rng(seed)
    pairs = []
    for i in range(n):
        kind = i % 3
        pa = rng.uniform(-20, 20, 2)
        pb = pa + rng.uniform(-1.6, 1.6, 2)

        def shape(is_poly, p):
            rot = float(np.float32(rng.uniform(-7, 7)))
            if is_poly:
                k = int(rng.integers(3, 9))
                ang = np.sort(rng.uniform(0, 2 * np.pi, k))
                r = rng.uniform(0.4, 1.2)
                v = np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)
                return dict(shape=1, verts=v, pos=(float(np.float32(p[0])), float(np.float32(p[1]))), rot=rot)
            w = rng.uniform(0.4, 2.5, 2).astype(np.float32)
            return dict(shape=0, width=(float(w[0]), float(w[1])), pos=(float(np.float32(p[0])), float(np.float32(p[1]))), rot=rot)

        pairs.append((shape(kind == 1, pa), shape(kind >= 1, pb)))
    return pairs


def mixed_scene():
    from sylt2d_b200 import scenes
    b = scenes._Builder()
    scenes._ground(b)
    b.polygon(scenes.HEXAGON, 5.0, 0.3, (-2.0, 1.2), 0.1)
    b.box(1.0, 1.0, 10.0, 0.2, (0.0, 0.6), 0.0)
    b.polygon(scenes.PENTAGON, 4.0, 0.4, (0.2, 2.3), -0.3, (0.0, -1.0), 0.5)
    b.box(1.5, 0.5, 6.0, 0.5, (2.0, 0.4), 0.05)
    b.polygon(scenes.regular_ngon(3, 0.8), 3.0, 0.2, (2.1, 1.6), 0.4)
    b.box(0.8, 0.8, 2.0, 0.3, (-2.1, 2.6), 0.7)
    return b.scene("mixed", iterations=10)


def step_cases():
    from sylt2d_b200 import scenes
    return [("pyramid5", scenes.pyramid(5, 10), 90), ("mixed", mixed_scene(), 80), ("bridge", scenes.bridge(10, omega0=0.5), 60),
            ("multi_pendulum", scenes.multi_pendulum(10, omega0=0.7), 60), ("demo2", scenes.demo2(10), 60)]


def run_oracle(scene, steps, mode=0):
    from oracle import orc
    o = orc.OracleWorld(scene.gravity, scene.iterations, sincos_mode=mode)
    o.load_scene(scene)
    for _ in range(steps):
        o.step(scene.dt)
    st = o.get_bodies()
    arb, con = o.get_arbiters()
    state = np.stack([st[f] for f in STATE_FIELDS], 1).astype(np.float32)
    keys = np.stack([arb["id1"], arb["id2"], arb["num_contacts"]], 1).astype(np.int64)
    imp = np.stack([con["pn"], con["pt"]], 1).astype(np.float32)
    return state, keys, imp


def main():
    from oracle import orc
    orc.build()
    pairs = narrow_inputs()
    counts, contacts = [], []
    for a, b in pairs:
        got = orc.collide_pair(a, b, sincos_mode=0)
        counts.append(len(got))
        contacts.append(np.stack([got[f].astype(np.float64) for f in CONTACT_FIELDS], 1) if len(got) else np.zeros((0, len(CONTACT_FIELDS))))
    con = np.concatenate(contacts).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "narrow_pairs.npz"), counts=np.asarray(counts, np.int32), contacts=con, seed=np.int32(11), n=np.int32(len(pairs)))
    out = {}
    for name, sc, steps in step_cases():
        state, keys, imp = run_oracle(sc, steps)
        out[name + "_state"], out[name + "_keys"], out[name + "_impulses"], out[name + "_steps"] = state, keys, imp, np.int32(steps)
    np.savez_compressed(os.path.join(HERE, "steps.npz"), **out)
    print("hits", int((np.asarray(counts) > 0).sum()), "of", len(counts), "; contacts", con.shape[0])


if __name__ == "__main__":
    main()
