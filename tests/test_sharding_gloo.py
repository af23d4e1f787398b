"""world_size-2 gloo test of the N>1 host logic: contiguous world shards, no data-path collective, and the final
stats reduction (the only cross-rank traffic of a run).  The per-world stepping is stood in for by the
CPU oracle here (no GPU in this container); the real multi-GPU path is exercised by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import orc
    from sylt2d_b200 import scenes
    from sylt2d_b200.batch import shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    sc = scenes.bridge(10)
    n = 5
    bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=3, omega_jitter=1.0)
    lo, hi = shard_range(n, rank, world_size)
    nb, nj = sc.num_bodies, sc.joints.shape[0]
    ke, steps = 0.0, 0
    finals = []
    for k in range(lo, hi):
        w = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1)
        w.load_scene(scenes.Scene(sc.name, bodies[k * nb:(k + 1) * nb], joints[k * nj:(k + 1) * nj], sc.verts, sc.gravity, sc.iterations, sc.dt, sc.ctx))
        for _ in range(20):
            w.step(sc.dt)
        b = w.get_bodies()
        finals.append((k, b))
        ke += float((b["vx"].astype(np.float64) ** 2 + b["vy"].astype(np.float64) ** 2).sum())
        steps += nb * 20
    t = torch.tensor([ke, float(steps)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the single final reduction
    q.put((rank, finals, t.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference
    q1 = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port1 = s.getsockname()[1]
    p1 = ctx.Process(target=_worker, args=(0, 1, port1, q1))
    p1.start()
    single = q1.get(timeout=120)
    p1.join(timeout=60)
    worlds = {}
    for _, finals, tot in res:
        for k, b in finals:
            assert k not in worlds
            worlds[k] = b
        assert abs(tot[0] - single[2][0]) < 1e-9 * max(1.0, single[2][0]) and tot[1] == single[2][1]
    assert sorted(worlds) == [0, 1, 2, 3, 4]
    for k, b in single[1]:
        assert worlds[k].tobytes() == b.tobytes()  # per-world results do not depend on the number of ranks
