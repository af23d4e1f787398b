"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Gates:
  * arbiter key set + per-key contact count after the broad phase: bit-exact vs the oracle's all-pairs loop
  * contacts (position / normal / separation / feature edges) from identical state: bit-exact (tolerance 0;
    the north star allows 1e-5)
  * single-step and multi-step solver output, oracle replaying the GPU's colour order: bit-exact (tolerance 0;
    the north star allows a stated tolerance)
  * colour order vs canonical (id1,id2) order: stated looser tolerance + long-run penetration / energy
"""
import os

import numpy as np
import pytest

from oracle import orc
from sylt2d_b200 import scenes

pytestmark = pytest.mark.gpu

STATE_FIELDS = ("x", "y", "rotation", "vx", "vy", "angular_velocity")
CONTACT_GEOM = ("px", "py", "nx", "ny", "separation", "in_edge_1", "out_edge_1", "in_edge_2", "out_edge_2")
CONTACT_SOLVE = ("pn", "pt", "pnb", "mass_normal", "mass_tangent", "bias", "r1x", "r1y", "r2x", "r2y")


def _mk(scene, **kw):
    from sylt2d_b200 import World
    w = World(scene.gravity, scene.iterations, device=0)
    w.load_scene(scene)
    o = orc.OracleWorld(scene.gravity, scene.iterations, sincos_mode=1)
    o.load_scene(scene)
    return w, o


def _assert_states_equal(g, r, what=""):
    for f in STATE_FIELDS:
        bad = np.nonzero(~((g[f] == r[f]) | (np.isnan(g[f]) & np.isnan(r[f]))))[0]
        assert bad.size == 0, f"{what}: field {f} differs at bodies {bad[:8]}: gpu {g[f][bad[:4]]} oracle {r[f][bad[:4]]}"


def _assert_arbiters_equal(w, o, solve_fields=True, what=""):
    ga, gc = w.get_arbiters()
    ra, rc = o.get_arbiters()
    assert ga.shape[0] == ra.shape[0], f"{what}: arbiter count {ga.shape[0]} vs oracle {ra.shape[0]}"
    assert np.array_equal(ga["id1"], ra["id1"]) and np.array_equal(ga["id2"], ra["id2"]), f"{what}: arbiter key sets differ"
    assert np.array_equal(ga["num_contacts"], ra["num_contacts"]), f"{what}: contact counts differ"
    assert np.array_equal(ga["friction"], ra["friction"]), f"{what}: friction differs"
    for f in CONTACT_GEOM + (CONTACT_SOLVE if solve_fields else ()):
        ok = (gc[f] == rc[f]) | (np.isnan(gc[f].astype(np.float64)) & np.isnan(rc[f].astype(np.float64)))
        assert ok.all(), f"{what}: contact field {f} differs at {np.nonzero(~ok)[0][:6]}: {gc[f][~ok][:4]} vs {rc[f][~ok][:4]}"
    return ga


def _run_lockstep(scene, steps, check_every=1, ctx=None):
    w, o = _mk(scene)
    if ctx is not None:
        from sylt2d_b200 import WorldContext
        w.world_context = WorldContext(*ctx)
        o.set_context(*ctx)
    for s in range(steps):
        w.step(scene.dt)
        keys, jo = w.solve_order()
        o.step_ordered(scene.dt, keys, jo)
        if s % check_every == 0 or s == steps - 1:
            _assert_states_equal(w.get_bodies(), o.get_bodies(), f"{scene.name} step {s}")
            ga = _assert_arbiters_equal(w, o, what=f"{scene.name} step {s}")
            # a proper colouring: no two arbiters of one colour share a dynamic body
            props = w.get_body_props()
            for c in np.unique(ga["color"]):
                sel = ga[ga["color"] == c]
                b = np.concatenate([sel["body1"], sel["body2"]])
                b = b[props[b, 1] != 0.0]
                assert np.unique(b).size == b.size, f"colour {c} is not an independent set"
    gj, rj = w.get_joints(), o.get_joints()
    for f in gj.dtype.names:
        assert np.array_equal(gj[f], rj[f]), f"{scene.name}: joint field {f} differs"
    return w, o


def test_device_sincos_matches_host_bitwise():
    import ctypes as C
    from sylt2d_b200 import device_sincos
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e4, 1e4, 100000), rng.uniform(-1e-3, 1e-3, 1000),
                        np.array([0.0, -0.0, 0.75, 0.785398, 119.99, 120.0, 121.0, 1e6, -1e6, 3.4e38, 1e-40, np.pi, -np.pi / 2])]).astype(np.float32)
    s, c = device_sincos(a)
    L = orc.lib()
    hs, hc = np.zeros_like(a), np.zeros_like(a)
    ss, cc = C.c_float(), C.c_float()
    for i in range(0, a.shape[0], 7):  # subsample for speed
        L.orc_sincos(1, float(a[i]), C.byref(ss), C.byref(cc))
        hs[i], hc[i] = ss.value, cc.value
    idx = np.arange(0, a.shape[0], 7)
    assert np.array_equal(s[idx].view(np.uint32), hs[idx].view(np.uint32))
    assert np.array_equal(c[idx].view(np.uint32), hc[idx].view(np.uint32))


def _desc(shape, w=(0, 0), pos=(0, 0), rot=0.0, voff=0, vcnt=0):
    d = np.zeros(1, dtype=scenes.BODY_DESC)
    d["id"], d["shape"], d["vert_offset"], d["vert_count"] = 1, shape, voff, vcnt
    d["width_x"], d["width_y"], d["mass"], d["x"], d["y"], d["rotation"] = w[0], w[1], 1.0, pos[0], pos[1], rot
    return d


def test_narrow_phase_boxes_bit_exact():
    from sylt2d_b200 import collide_pairs
    rng = np.random.default_rng(3)
    n = 4000
    A = np.zeros(n, dtype=scenes.BODY_DESC)
    B = np.zeros(n, dtype=scenes.BODY_DESC)
    for D in (A, B):
        D["mass"] = 1.0
        D["width_x"], D["width_y"] = rng.uniform(0.3, 3.0, n), rng.uniform(0.3, 3.0, n)
        D["rotation"] = rng.uniform(-7, 7, n)
    A["x"], A["y"] = rng.uniform(-50, 50, n), rng.uniform(-50, 50, n)
    B["x"], B["y"] = A["x"] + rng.uniform(-2.5, 2.5, n), A["y"] + rng.uniform(-2.5, 2.5, n)
    out, cnt = collide_pairs(A, B, cap=4)
    hits = 0
    for i in range(n):
        exp = orc.collide_pair(dict(shape=0, width=(A["width_x"][i], A["width_y"][i]), pos=(A["x"][i], A["y"][i]), rot=A["rotation"][i]),
                               dict(shape=0, width=(B["width_x"][i], B["width_y"][i]), pos=(B["x"][i], B["y"][i]), rot=B["rotation"][i]),
                               sincos_mode=1)
        assert cnt[i] == len(exp), i
        for f in CONTACT_GEOM:
            assert np.array_equal(out[i, :cnt[i]][f], exp[f]), (i, f)
        hits += cnt[i] > 0
    assert hits > 500


def test_narrow_phase_polygons_bit_exact():
    from sylt2d_b200 import collide_pairs
    rng = np.random.default_rng(5)
    n = 1500
    verts, A, B = [], np.zeros(n, dtype=scenes.BODY_DESC), np.zeros(n, dtype=scenes.BODY_DESC)
    nv = 0
    descs = []
    for i in range(n):
        pa = rng.uniform(-20, 20, 2)
        pb = pa + rng.uniform(-1.5, 1.5, 2)
        pair = []
        for D, p, force_poly in ((A, pa, True), (B, pb, i % 3 != 0)):
            D["mass"][i], D["x"][i], D["y"][i], D["rotation"][i] = 1.0, p[0], p[1], rng.uniform(-7, 7)
            if force_poly:
                k = int(rng.integers(3, 9))
                ang = np.sort(rng.uniform(0, 2 * np.pi, k))
                r = rng.uniform(0.4, 1.2)
                v = np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)
                D["shape"][i], D["vert_offset"][i], D["vert_count"][i] = 1, nv, k
                verts.append(v)
                nv += k
                pair.append(dict(shape=1, verts=v, pos=(D["x"][i], D["y"][i]), rot=D["rotation"][i]))
            else:
                D["width_x"][i], D["width_y"][i] = rng.uniform(0.4, 2.0, 2)
                pair.append(dict(shape=0, width=(D["width_x"][i], D["width_y"][i]), pos=(D["x"][i], D["y"][i]), rot=D["rotation"][i]))
        descs.append(pair)
    V = np.concatenate(verts)
    out, cnt = collide_pairs(A, B, V, cap=16)
    hits = 0
    for i in range(n):
        exp = orc.collide_pair(descs[i][0], descs[i][1], sincos_mode=1)
        assert cnt[i] == len(exp), i
        for f in CONTACT_GEOM:
            assert np.array_equal(out[i, :cnt[i]][f], exp[f]), (i, f)
        hits += cnt[i] > 0
    assert hits > 300


def test_polygons_up_to_32_vertices_bit_exact(monkeypatch):
    """The narrow phase comes in tiers of 8 / 16 / 32 vertices per polygon (the reference is unbounded, src/body.rs:246): pairs of
    9..32-gons through sylt_collide_pairs, worlds whose largest polygon has 12, 20 and 32 vertices in lock-step with the oracle (both
    single-world solvers), and a 33-gon is refused."""
    from sylt2d_b200 import collide_pairs, World, Body, SyltError
    rng = np.random.default_rng(9)
    n = 300
    verts, A, B = [], np.zeros(n, dtype=scenes.BODY_DESC), np.zeros(n, dtype=scenes.BODY_DESC)
    nv, descs = 0, []
    for i in range(n):
        pa = rng.uniform(-20, 20, 2)
        pb = pa + rng.uniform(-1.2, 1.2, 2)
        pair = []
        for D, p in ((A, pa), (B, pb)):
            k = int(rng.integers(9, 33))
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            r = rng.uniform(0.5, 1.2)
            v = np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)
            D["mass"][i], D["x"][i], D["y"][i], D["rotation"][i] = 1.0, p[0], p[1], rng.uniform(-7, 7)
            D["shape"][i], D["vert_offset"][i], D["vert_count"][i] = 1, nv, k
            verts.append(v)
            nv += k
            pair.append(dict(shape=1, verts=v, pos=(D["x"][i], D["y"][i]), rot=D["rotation"][i]))
        descs.append(pair)
    out, cnt = collide_pairs(A, B, np.concatenate(verts), cap=64)
    hits = 0
    for i in range(n):
        exp = orc.collide_pair(descs[i][0], descs[i][1], sincos_mode=1)
        assert cnt[i] == len(exp), i
        for f in CONTACT_GEOM:
            assert np.array_equal(out[i, :cnt[i]][f], exp[f]), (i, f)
        hits += cnt[i] > 2
    assert hits > 50  # manifolds with more contacts than a box pair can have
    for regions in ("0", "2"):
        monkeypatch.setenv("SYLT_REGIONS", regions)
        for k in (12, 20, 32):
            b = scenes._Builder()
            scenes._ground(b)
            for i in range(6):
                b.polygon(scenes.regular_ngon(k if i % 2 == 0 else 5, 0.6, 0.1 * i), 4.0, 0.4, (-4.0 + 1.5 * i, 1.0 + 0.9 * (i % 3)), 0.2 * i)
                b.box(0.8, 0.8, 3.0, 0.4, (-4.0 + 1.5 * i, 4.0 + 0.5 * (i % 2)))
            _run_lockstep(b.scene(f"ngon{k}", iterations=10), 50, check_every=5)
    w = World((0.0, -10.0), 10, device=0)
    with pytest.raises(SyltError):
        w.add_body(Body.new_polygon(scenes.regular_ngon(33, 1.0), 1.0))


@pytest.mark.parametrize("name,steps", [("pyramid6", 120), ("demo1", 150), ("demo2", 80), ("demo3", 150), ("demo4", 120), ("demo6", 150),
                                        ("demo8", 120), ("demo10", 100), ("bridge", 100), ("multi_pendulum", 100)])
def test_world_step_bit_exact_in_gpu_order(name, steps):
    sc = scenes.pyramid(6, 10) if name == "pyramid6" else getattr(scenes, name)(10)
    _run_lockstep(sc, steps)


@pytest.mark.parametrize("ctx", [(a, w, p) for a in (True, False) for w in (True, False) for p in (True, False)])
def test_world_context_flag_combinations(ctx):
    # quirks Q-4 / Q-5: three independent flags, all 8 combinations
    sc = scenes.demo6(10)
    _run_lockstep(sc, 60, ctx=ctx)
    _run_lockstep(scenes.pyramid(4, 10), 60, ctx=ctx)


def test_reference_literal_pyramid_12_rows_100_iterations():
    # the reference's own demo5 settings (quirk Q-14)
    sc = scenes.pyramid(12, 100)
    _run_lockstep(sc, 40, check_every=10)


def _random_scene(n, seed, polys=True, span=30.0):
    rng = np.random.default_rng(seed)
    b = scenes._Builder()
    b.box(2 * span + 10, 2.0, scenes.F32_MAX, 0.4, (0.0, -1.0))
    b.box(2.0, 40.0, scenes.F32_MAX, 0.4, (-span - 4.0, 18.0))
    b.box(2.0, 40.0, scenes.F32_MAX, 0.4, (span + 4.0, 18.0))
    for _ in range(n):
        pos = (rng.uniform(-span, span), rng.uniform(0.5, 12.0))
        rot = rng.uniform(-np.pi, np.pi)
        if polys and rng.random() < 0.5:
            k = int(rng.integers(3, 9))
            b.polygon(scenes.regular_ngon(k, rng.uniform(0.3, 0.8)), rng.uniform(1, 20), rng.uniform(0.1, 0.6), pos, rot,
                      vel=(rng.uniform(-2, 2), rng.uniform(-2, 2)), omega=rng.uniform(-3, 3))
        else:
            b.box(rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(1, 20), rng.uniform(0.1, 0.6), pos, rot,
                  vel=(rng.uniform(-2, 2), rng.uniform(-2, 2)), omega=rng.uniform(-3, 3))
    return b.scene(f"random{n}_{seed}", iterations=10)


@pytest.mark.parametrize("seed,polys", [(1, False), (2, True), (3, True)])
def test_broad_phase_pair_set_bit_exact_vs_all_pairs(seed, polys):
    # grid broad phase (GPU) vs the reference's all-pairs loop (oracle): same arbiter keys, same contacts
    sc = _random_scene(700, seed, polys)
    w, o = _mk(sc)
    w.broad_phase()
    o.broad_phase()
    ga = _assert_arbiters_equal(w, o, solve_fields=False, what=sc.name)
    assert ga.shape[0] > 300
    st = w.get_stats()
    assert st["candidate_pairs"] >= st["arbiters"]


@pytest.mark.parametrize("seed,polys", [(11, False), (12, True)])
def test_random_pile_steps_bit_exact(seed, polys):
    sc = _random_scene(400, seed, polys)
    _run_lockstep(sc, 40, check_every=5)


def test_body_with_more_than_32_neighbours_uses_wide_colour_mask():
    """A dynamic plank carrying 45 boxes needs 46 colours (one per arbiter on the plank); bit-exact vs the oracle."""
    b = scenes._Builder()
    scenes._ground(b)
    b.box(60.0, 1.0, 500.0, 0.3, (0.0, 0.55))
    for i in range(45):
        b.box(1.0, 1.0, 5.0, 0.3, (-27.5 + 1.25 * i, 1.6))
    sc = b.scene("plank45", iterations=10)
    w, o = _run_lockstep(sc, 60, check_every=10)
    assert w.get_stats()["colors"] >= 46
    ga, _ = w.get_arbiters()
    assert ga["color"].max() >= 32


def test_bodies_added_between_steps_and_clear():
    # launch_bomb (examples/samples/src/main.rs:56-64): add_body between steps; World::clear + reload
    from sylt2d_b200 import Body, Vec2
    sc = scenes.pyramid(5, 10)
    w, o = _mk(sc)
    nid = int(sc.bodies["id"].max()) + 1
    for s in range(60):
        if s in (10, 25):
            bomb = Body.new(Vec2(1.0, 1.0), 50.0)
            bomb.id = nid
            bomb.friction = 0.2
            bomb.position = Vec2(-3.0 + 0.3 * s, 15.0)
            bomb.rotation = 0.3
            bomb.velocity = Vec2(bomb.position.x * -1.5, 15.0 * -1.5)
            bomb.angular_velocity = 7.0
            w.add_body(bomb)
            o.add_box(nid, (1.0, 1.0), 50.0, 0.2, tuple(bomb.position), 0.3, tuple(bomb.velocity), 7.0)
            nid += 1
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"bomb step {s}")
    _assert_arbiters_equal(w, o, what="bomb")
    w.clear()
    o.clear()
    assert w.num_bodies == 0
    sc2 = scenes.demo3(10)
    w.load_scene(sc2)
    o.load_scene(sc2)
    for s in range(30):
        w.step(sc2.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc2.dt, keys, jo)
    _assert_states_equal(w.get_bodies(), o.get_bodies(), "after clear")


def test_mutating_pub_fields_between_steps():
    sc = scenes.demo4(10)
    w, o = _mk(sc)
    for s in range(40):
        if s == 15:
            st = w.get_bodies()
            st["vx"][3] = 4.0
            st["angular_velocity"][5] = -2.0
            w.set_bodies(st)
            o.set_bodies(st)
            w.add_force(2, (30.0, 5.0))
            o.add_force(2, (30.0, 5.0))
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"mutate step {s}")


def test_errors_match_reference_contract():
    from sylt2d_b200 import World, Body, Joint, Vec2, SyltError
    # NoInverse mid-step (quirk Q-15): joint between two static bodies
    w = World((0, -10), 10)
    a, b = Body.new(Vec2(1, 1), scenes.F32_MAX), Body.new(Vec2(1, 1), scenes.F32_MAX)
    b.position = Vec2(3, 0)
    w.add_body(a)
    w.add_body(b)
    w.add_joint(Joint.new(a, b, Vec2(1.5, 0.0), w))
    with pytest.raises(SyltError) as ei:
        w.step(1 / 60)
    assert ei.value.code == 1
    j, m = w.last_error_matrix()
    assert j == 0 and list(m) == [0, 0, 0, 0]
    # out-of-order ids (quirk Q-6): the reference panics, the ABI returns SYLT_ERR_BODY_ORDER
    w2 = World((0, -10), 10)
    c, d = Body.new(Vec2(1, 1), 1.0), Body.new(Vec2(1, 1), 1.0)
    w2.add_body(d)
    w2.add_body(c)
    with pytest.raises(SyltError) as ei:
        w2.step(1 / 60)
    assert ei.value.code == 3
    # Joint::new on a body that is not in the world (joint.rs:31 `expect`)
    with pytest.raises(SyltError) as ei:
        Joint.new(a, c, Vec2(0, 0), w)
    assert ei.value.code == 4
    # empty world steps fine
    w3 = World((0, -10), 10)
    w3.step(1 / 60)
    assert w3.num_arbiters == 0


def test_step_n_equals_n_steps_and_is_deterministic():
    sc = scenes.pyramid(8, 10)
    from sylt2d_b200 import World
    res = []
    for mode in range(3):
        w = World(sc.gravity, sc.iterations)
        w.load_scene(sc)
        if mode == 0:
            for _ in range(50):
                w.step(sc.dt)
        elif mode == 1:
            w.step_n(sc.dt, 50)
        else:
            w.step_n_async(sc.dt, 20)
            w.step_n_async(sc.dt, 30)
            w.sync()
        res.append(w.get_bodies())
    _assert_states_equal(res[0], res[1], "step vs step_n")
    _assert_states_equal(res[0], res[2], "step vs async")


def test_colour_order_vs_canonical_order_tolerance_and_long_run():
    """The reference's own arbiter order is a randomised HashMap order (quirk Q-1), so any order is "the reference";
    canonical ascending-key order is what BASELINE.md calls reference keying.  Single step: |dv| <= 5e-2 m/s on a
    settled 20-row pyramid under g=10 (stated looser tolerance).  Long run: penetration and energy stay in family."""
    sc = scenes.pyramid(20, 10)
    w, o = _mk(sc)
    for s in range(300):
        w.step(sc.dt)
    o_can = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1)
    o_can.load_scene(sc)
    for s in range(300):
        o_can.step(sc.dt)
    g, r = w.get_bodies(), o_can.get_bodies()
    # Which boxes topple off the top while the pyramid lands is chaotic in the solve order (the reference's own order
    # is random), so the long-run gate is on robust statistics: almost every body at rest, same stack height for the
    # resting bodies, bounded penetration, similar contact count.
    sp_g = np.sqrt(g["vx"] ** 2 + g["vy"] ** 2)
    sp_r = np.sqrt(r["vx"] ** 2 + r["vy"] ** 2)
    assert (sp_g < 0.1).mean() > 0.97 and (sp_r < 0.1).mean() > 0.97, ((sp_g < 0.1).mean(), (sp_r < 0.1).mean())
    assert abs(np.median(g["y"]) - np.median(r["y"])) < 0.05
    assert abs(np.percentile(g["y"], 95) - np.percentile(r["y"], 95)) < 0.5
    _, gc = w.get_arbiters()
    _, rc = o_can.get_arbiters()
    assert np.percentile(gc["separation"], 1) > -0.05 and np.percentile(rc["separation"], 1) > -0.05
    assert abs(gc.shape[0] - rc.shape[0]) <= 0.1 * rc.shape[0]
    # single step from identical state, different order
    st = o_can.get_bodies()
    w2, o2 = _mk(sc)
    w2.set_bodies(st)
    o2.set_bodies(st)
    w2.step(sc.dt)
    o2.step(sc.dt)  # canonical order
    a, b = w2.get_bodies(), o2.get_bodies()
    assert np.abs(a["vx"] - b["vx"]).max() < 5e-2 and np.abs(a["vy"] - b["vy"]).max() < 5e-2


# ------------------------------------------------------------------ batched worlds
def _batch_lockstep(scene, n_worlds, steps, x_jitter=0.0, omega_jitter=0.0, check_worlds=(0, 1), capacity=None):
    from sylt2d_b200 import Batch
    bodies, boff, joints, joff = scenes.tile_worlds(scene, n_worlds, seed=7, x_jitter=x_jitter, omega_jitter=omega_jitter)
    b = Batch(bodies, boff, joints, joff, scene.gravity, scene.iterations, device=0, ctx=scene.ctx)
    if capacity is not None:
        b.set_capacity(capacity)
    nb = scene.num_bodies
    oracles = {}
    for wi in check_worlds:
        sc = scenes.Scene(scene.name, bodies[wi * nb:(wi + 1) * nb], scene.joints, scene.verts, scene.gravity, scene.iterations, scene.dt, scene.ctx)
        o = orc.OracleWorld(scene.gravity, scene.iterations, sincos_mode=1)
        o.load_scene(sc)
        oracles[wi] = o
    for s in range(steps):
        b.step_n(scene.dt, 1)
        for wi, o in oracles.items():
            keys, jo = b.solve_order(wi)
            o.step_ordered(scene.dt, keys, jo)
            _assert_states_equal(b.get_bodies(wi, 1), o.get_bodies(), f"batch {scene.name} world {wi} step {s}")
    for wi, o in oracles.items():
        ga, gc = b.get_arbiters(wi)
        ra, rc = o.get_arbiters()
        assert np.array_equal(ga["id1"], ra["id1"]) and np.array_equal(ga["id2"], ra["id2"])
        for f in CONTACT_GEOM + CONTACT_SOLVE:
            assert np.array_equal(gc[f], rc[f]), f
        gj, rj = b.get_joints(wi), o.get_joints()
        for f in gj.dtype.names:
            assert np.array_equal(gj[f], rj[f]), f
    assert not b.get_errors().any()
    return b


def test_batch_pyramid_bit_exact():
    _batch_lockstep(scenes.pyramid(8, 10), 6, 80, x_jitter=0.01, check_worlds=(0, 5))


def test_batch_pyramid20_config5_world():
    _batch_lockstep(scenes.pyramid(20, 10), 3, 40, x_jitter=0.01, check_worlds=(1,))


def test_batch_large_world_colour_wider_than_the_cta():
    """785 bodies in 28 stacks: two colours of ~380 arbiters each (more than the CTA has threads), slots span several rounds."""
    bd = scenes._Builder()
    scenes._ground(bd)
    for c in range(28):
        for r in range(28):  # slightly overlapping vertically: every box touches the one below from the first step
            bd.box(1.0, 1.0, 10.0, 0.2, (-14.0 + 1.05 * c, 0.49 + 0.98 * r))
    b = _batch_lockstep(bd.scene("stacks28", iterations=10), 2, 12, x_jitter=0.002, check_worlds=(1,))
    arb, _ = b.get_arbiters(1, cap=8192)
    assert arb.shape[0] > 700 and np.bincount(arb["color"]).max() > 256


def test_batch_slot_capacity():
    """A dense block has ~4 candidate pairs per body: over the default capacity (SYLT_ERR_CAPACITY, world frozen) until raised."""
    from sylt2d_b200 import Batch, SyltError
    bd = scenes._Builder()
    scenes._ground(bd)
    for r in range(12):
        for c in range(12):
            bd.box(1.0, 1.0, 10.0, 0.2, (-6.0 + 0.98 * c, 0.49 + 0.98 * r))
    sc = bd.scene("block12", iterations=10)
    small = Batch(*scenes.tile_worlds(sc, 2, seed=7, x_jitter=0.002), sc.gravity, sc.iterations, device=0, ctx=sc.ctx)
    small.step_n(sc.dt, 1)
    assert (small.get_errors() == 5).all()
    with pytest.raises(SyltError):
        small.set_capacity(60000)  # no longer fits the shared memory of one CTA
    b = _batch_lockstep(sc, 2, 12, x_jitter=0.002, check_worlds=(0,), capacity=700)
    arb, _ = b.get_arbiters(0)
    assert arb.shape[0] > 250


def test_batch_bridge_and_pendulum_bit_exact():
    _batch_lockstep(scenes.bridge(10), 5, 80, omega_jitter=1.0, check_worlds=(0, 4))
    _batch_lockstep(scenes.multi_pendulum(10), 5, 80, omega_jitter=1.0, check_worlds=(0, 3))
    _batch_lockstep(scenes.demo8(10), 2, 60, check_worlds=(1,))


def test_batch_multi_step_launch_equals_single_steps_and_shards():
    # world results must not depend on how worlds are sharded over GPUs nor on steps per launch
    from sylt2d_b200 import Batch, shard_range
    sc = scenes.pyramid(6, 10)
    n = 9
    bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=3, x_jitter=0.01)
    full = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
    full.step_n(sc.dt, 45)
    ref = full.get_bodies()
    nb = sc.num_bodies
    for ws in (2, 4):
        parts = []
        for r in range(ws):
            lo, hi = shard_range(n, r, ws)
            if hi <= lo:
                continue
            sb = Batch(bodies[lo * nb:hi * nb], boff[lo:hi + 1] - boff[lo], joints[:0], joff[lo:hi + 1] - joff[lo], sc.gravity, sc.iterations, ctx=sc.ctx)
            for _ in range(45):
                sb.step_n(sc.dt, 1)
            parts.append(sb.get_bodies())
        _assert_states_equal(np.concatenate(parts), ref, f"shards={ws}")
    st = full.stats()
    assert st["body_steps"] == n * nb * 45 and st["errors"] == 0 and st["contacts"] > 0


# ------------------------------------------------------------------ BASELINE.json full-size configurations
def _lockstep_big(sc, settle, steps):
    from sylt2d_b200 import World
    w = World(sc.gravity, sc.iterations, device=0)
    w.load_scene(sc)
    w.step_n(sc.dt, settle)                       # get into the interesting regime on the GPU ...
    o = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1, broadphase=1)   # grid candidates == all-pairs set (CPU test)
    o.load_scene(sc)
    st = w.get_bodies()
    o.set_bodies(st)                              # ... then hand the same state to both and compare from there
    w2 = World(sc.gravity, sc.iterations, device=0)
    w2.load_scene(sc)
    w2.set_bodies(st)
    for s in range(steps):
        w2.step(sc.dt)
        keys, jo = w2.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w2.get_bodies(), o.get_bodies(), f"{sc.name} step {s}")
    _assert_arbiters_equal(w2, o, what=sc.name)
    return w2


def test_config2_mixed_pile_10k_bit_exact():
    sc = scenes.mixed_pile(10000, seed=1234)
    w = _lockstep_big(sc, 150, 4)
    st = w.get_stats()
    assert st["arbiters"] > 3000 and st["contacts"] > st["arbiters"]


def test_config3_box_drop_100k_bit_exact():
    sc = scenes.box_drop(100000, seed=42)
    w = _lockstep_big(sc, 200, 3)
    st = w.get_stats()
    assert st["bodies"] == 100003 and st["arbiters"] > 50000


def test_config4_joint_worlds_batch_properties():
    # 8192 joint worlds (4096 bridge + 4096 multi-pendulum): size-independent properties at full size —
    # no errors, finite state, planks stay attached (bounded distance from their anchors), identical worlds stay identical
    from sylt2d_b200 import Batch
    for sc, jit in ((scenes.bridge(10), 0.0), (scenes.multi_pendulum(10), 1.0)):
        n = 4096
        bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=5, omega_jitter=jit)
        b = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
        b.step_n(sc.dt, 120)
        st = b.get_bodies()
        assert not b.get_errors().any()
        assert np.isfinite(st["x"]).all() and np.isfinite(st["y"]).all()
        nb = sc.num_bodies
        if jit == 0.0:  # unperturbed copies must be bit-identical across the batch
            first = st[:nb].tobytes()
            for k in (1, 777, n - 1):
                assert st[k * nb:(k + 1) * nb].tobytes() == first
        assert np.abs(st["x"]).max() < 60 and st["y"].min() > -30


def test_batch_step_host_equals_set_step_get():
    from sylt2d_b200 import Batch
    sc = scenes.pyramid(6, 10)
    n = 1100   # > 2 pipeline chunks
    bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=9, x_jitter=0.01)
    a = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
    b = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
    a.step_n(sc.dt, 30)
    b.step_n(sc.dt, 30)
    st = a.get_bodies()
    st["vx"][5::7] += 0.25          # the caller mutated some state on the host
    out = np.zeros_like(st)
    for _ in range(3):
        a.set_bodies(st)
        a.step_n(sc.dt, 1)
        ref = a.get_bodies()
        b.step_host(sc.dt, 1, st, out)
        _assert_states_equal(out, ref, "step_host")
        st = ref.copy()


def test_cpp_mirror_and_headless_runner(tmp_path):
    import subprocess
    import sys
    from tests.test_abi import _build_cpp_example, ROOT
    import __graft_entry__ as g
    g.build()
    exe = _build_cpp_example(str(tmp_path))
    r = subprocess.run([exe, "300"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr          # the reference's demo5 settles to an ~11.4 high pyramid
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_demo.py"), "demo8", "--steps", "120", "--bombs", "2", "--iterations", "10"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "demo8" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_api_sequences_vs_oracle(seed):
    """Randomised caller behaviour between steps (what examples/samples does through the GUI): add boxes / polygons /
    joints, push forces, overwrite body state, toggle the three WorldContext flags, change dt (including dt <= 0, the
    sample's Left-arrow key: inv_dt = 0, src/world.rs:108) — always bit-identical to the oracle replaying the GPU order."""
    from sylt2d_b200 import World, WorldContext, Body, Joint, Vec2
    rng = np.random.default_rng(100 + seed)
    w = World((0.0, -10.0), 10)
    o = orc.OracleWorld((0.0, -10.0), 10, sincos_mode=1)
    ids = []
    nid = 1

    def add_box(dynamic=True):
        nonlocal nid
        b = Body.new(Vec2(float(rng.uniform(0.4, 1.6)), float(rng.uniform(0.4, 1.6))), float(rng.uniform(10, 30)) if dynamic else scenes.F32_MAX)
        b.id = nid
        b.friction = float(rng.uniform(0.0, 0.7))
        b.position = Vec2(float(rng.uniform(-6, 6)), float(rng.uniform(0.5, 9.0)))
        b.rotation = float(rng.uniform(-3, 3))
        b.velocity = Vec2(float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)))
        b.angular_velocity = float(rng.uniform(-2, 2))
        w.add_body(b)
        o.add_box(nid, tuple(b.width), b.mass, b.friction, tuple(b.position), b.rotation, tuple(b.velocity), b.angular_velocity)
        ids.append(b)
        nid += 1

    def add_poly():
        nonlocal nid
        v = scenes.regular_ngon(int(rng.integers(3, 9)), float(rng.uniform(0.3, 0.9)))
        b = Body.new_polygon(v, float(rng.uniform(10, 30)))
        b.id = nid
        b.friction = float(rng.uniform(0.0, 0.7))
        b.position = Vec2(float(rng.uniform(-6, 6)), float(rng.uniform(0.5, 9.0)))
        b.rotation = float(rng.uniform(-3, 3))
        w.add_body(b)
        o.add_polygon(nid, v, b.mass, b.friction, tuple(b.position), b.rotation)
        ids.append(b)
        nid += 1

    ground = Body.new(Vec2(40.0, 2.0), scenes.F32_MAX)
    ground.id = nid
    ground.friction = 0.4
    ground.position = Vec2(0.0, -1.0)
    w.add_body(ground)
    o.add_box(nid, (40.0, 2.0), scenes.F32_MAX, 0.4, (0.0, -1.0))
    ids.append(ground)
    nid += 1
    for _ in range(12):
        add_box()
    ctx = [True, True, True]
    for s in range(90):
        r = rng.random()
        if r < 0.10:
            add_box()
        elif r < 0.18:
            add_poly()
        elif r < 0.26 and len(ids) > 3:
            # joint between a dynamic body and its nearest neighbour, anchored half-way: with the reference's det-scaled
            # "inverse" (quirk Q-3) long lever arms or light bodies blow up to inf/NaN within a few steps
            st = w.get_bodies()
            a = int(rng.integers(1, len(ids)))
            d2 = (st["x"] - st["x"][a]) ** 2 + (st["y"] - st["y"][a]) ** 2
            d2[a] = np.inf
            d2[0] = np.inf
            b = int(np.argmin(d2))
            if d2[b] < 9.0:
                anchor = Vec2(float(0.5 * (st["x"][a] + st["x"][b])), float(0.5 * (st["y"][a] + st["y"][b])))
                j = Joint.new(ids[a], ids[b], anchor, w)
                j.softness, j.bias_factor = float(rng.uniform(0, 0.05)), float(rng.uniform(0.05, 0.3))
                w.add_joint(j)
                o.add_joint(ids[a].id, ids[b].id, tuple(anchor), j.softness, j.bias_factor)
        elif r < 0.34:
            k = int(rng.integers(1, len(ids)))
            f = (float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50)))
            w.add_force(k, f)
            o.add_force(k, f)
        elif r < 0.42:
            st = w.get_bodies()
            k = int(rng.integers(1, len(ids)))
            st["vx"][k] += np.float32(rng.uniform(-3, 3))
            st["rotation"][k] += np.float32(rng.uniform(-0.5, 0.5))
            w.set_bodies(st)
            o.set_bodies(st)
        elif r < 0.50:
            ctx[int(rng.integers(0, 3))] ^= True
            w.world_context = WorldContext(*ctx)
            o.set_context(*ctx)
        dt = float(np.float32(1.0 / 60.0)) if rng.random() > 0.08 else float(rng.choice([0.0, -1.0 / 60.0, 1.0 / 120.0]))
        w.step(dt)
        keys, jo = w.solve_order()
        o.step_ordered(dt, keys, jo)
        ob = o.get_bodies()
        # Known, documented difference: once impulses are non-finite the reference poisons STATIC bodies too
        # (p * 0 = NaN) while static bodies are never written here; a world in that state is garbage either way.
        assert np.isfinite(ob["x"]).all() and np.isfinite(ob["vx"]).all(), "test scenario must stay finite"
        _assert_states_equal(w.get_bodies(), ob, f"seq {seed} step {s}")
    _assert_arbiters_equal(w, o, what=f"seq {seed}")
    gj, rj = w.get_joints(), o.get_joints()
    for f in gj.dtype.names:
        assert np.array_equal(gj[f], rj[f]) or (np.isnan(gj[f].astype(np.float64)) == np.isnan(rj[f].astype(np.float64))).all(), f


@pytest.mark.parametrize("fixes", [(True, False), (False, True), (True, True)])
def test_opt_in_corrected_physics_matches_oracle(fixes):
    """feature-id contact matching and the true Mat2x2 inverse behind flags (default off).  With the
    flags on the GPU must equal the oracle with the same flags, bit for bit; and the fixes must actually change physics."""
    from sylt2d_b200 import World, Batch
    for sc in (scenes.pyramid(8, 10), scenes.bridge(10), scenes.demo8(10), scenes.demo1(10)):
        w = World(sc.gravity, sc.iterations)
        w.load_scene(sc)
        w.set_fixes(*fixes)
        o = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1)
        o.load_scene(sc)
        o.set_fixes(*fixes)
        for s in range(80):
            w.step(sc.dt)
            keys, jo = w.solve_order()
            o.step_ordered(sc.dt, keys, jo)
            _assert_states_equal(w.get_bodies(), o.get_bodies(), f"fixes {fixes} {sc.name} step {s}")
        _assert_arbiters_equal(w, o, what=f"fixes {fixes} {sc.name}")
    if fixes[1]:
        sc = scenes.bridge(10)
        stiff, soft = World(sc.gravity, sc.iterations), World(sc.gravity, sc.iterations)
        for ww in (stiff, soft):
            ww.load_scene(sc)
        stiff.set_fixes(False, True)
        stiff.step_n(sc.dt, 300)
        soft.step_n(sc.dt, 300)
        ys, yr = stiff.get_bodies()["y"][1:], soft.get_bodies()["y"][1:]
        assert np.ptp(ys) < 0.05 and abs(ys.mean() - 5.0) < 0.5       # planks hang from their pins
        assert np.ptp(yr) > 1.0                                        # reference: det-scaled "inverse", very soft / unstable at 10 iterations
        bodies, boff, joints, joff = scenes.tile_worlds(sc, 3)
        b = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
        b.set_fixes(False, True)
        b.step_n(sc.dt, 300)
        _assert_states_equal(b.get_bodies(1, 1), stiff.get_bodies(), "batch true inverse")


def test_batch_step_host6_equals_step_host():
    from sylt2d_b200 import Batch
    sc = scenes.pyramid(5, 10)
    n = 1300
    bodies, boff, joints, joff = scenes.tile_worlds(sc, n, seed=2, x_jitter=0.01)
    a = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
    b = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, ctx=sc.ctx)
    a.step_n(sc.dt, 20)
    b.step_n(sc.dt, 20)
    st = a.get_bodies()
    names = ("x", "y", "rotation", "vx", "vy", "angular_velocity")
    s6 = np.ascontiguousarray(np.stack([st[f] for f in names], axis=1))
    out9, out6 = np.zeros_like(st), np.zeros_like(s6)
    for _ in range(3):
        a.step_host(sc.dt, 1, st, out9)
        b.step_host6(sc.dt, 1, s6, out6)
        for k, f in enumerate(names):
            assert np.array_equal(out9[f], out6[:, k]), f
        st, s6 = out9.copy(), out6.copy()


# ------------------------------------------------------------------ experimental region solver (k_solve_regions, opt-in)
@pytest.mark.parametrize("count,period", [(5, 3), (0, 4)])
def test_region_solver_bit_exact(monkeypatch, count, period):
    """SYLT_REGIONS=2 routes joint-free worlds through k_solve_regions (bodies partitioned into regions that live in shared
    memory, cross-region arbiters replayed redundantly from ghost bodies, device-side re-cut every `period` steps): the
    result must still be bit-identical to the oracle replaying the reported colour order.  count = regions (0: one per SM)."""
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGIONS_COUNT", str(count))
    monkeypatch.setenv("SYLT_REGIONS_PERIOD", str(period))
    w, _ = _run_lockstep(scenes.pyramid(7, 10), 90, check_every=3)
    ga, _ = w.get_arbiters()
    assert (ga["color"] < 32).any() and (ga["color"] >= 32).any()  # both classes occur: interior (e.g. box on the static ground) and cross-region
    _run_lockstep(scenes.demo8(10), 60, check_every=3)                               # polygons
    _run_lockstep(scenes.demo4(10), 60, check_every=3, ctx=(False, True, True))      # no accumulation
    w = _lockstep_big(scenes.mixed_pile(3000, seed=7), 120, 3)                       # boxes + polygons, a few thousand bodies
    assert w.get_stats()["arbiters"] > 1000
    _run_lockstep(scenes.bridge(10), 20)                                             # joints: falls back to k_solve


@pytest.mark.parametrize("count,local", [(1, 1), (3, 1), (3, 0)])
def test_region_solver_interior_colours_overflow_into_generic_ones(monkeypatch, count, local):
    """A dynamic plank carrying 45 boxes inside ONE region has more arbiters than there are interior colours (32): the region's CTA
    colours what it can by itself (block barriers, shared-memory claims) and promotes the rest to generic colours in grid-wide rounds
    (local = 0: the grid-wide rounds colour everything).  Bit-exact vs the oracle; colours stay proper."""
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGIONS_COUNT", str(count))
    monkeypatch.setenv("SYLT_COLOR_LOCAL", str(local))
    b = scenes._Builder()
    scenes._ground(b)
    b.box(60.0, 1.0, 500.0, 0.3, (0.0, 0.55))
    for i in range(45):
        b.box(1.0, 1.0, 5.0, 0.3, (-27.5 + 1.25 * i, 1.6))
    w, _ = _run_lockstep(b.scene("plank45", iterations=10), 40, check_every=5)
    ga, _ = w.get_arbiters()
    assert ga.shape[0] >= 46 and ga["color"].max() >= 32
    if count == 1:  # everything is interior to the one region: all 32 interior colours are taken at the plank, the rest was promoted
        assert (ga["color"] < 32).sum() >= 32


def test_region_solver_body_made_static_and_dynamic_again(monkeypatch):
    """A body whose inv_mass / inv_moi are zeroed between steps (and restored later) changes the class of its arbiters — an arbiter whose
    other body lives in another region is interior while the body is static and cross-region once it moves: the colours of such
    arbiters are dropped instead of kept (an interior colour at a dynamic body of another region would never be exported)."""
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGIONS_COUNT", "6")
    sc = scenes.pyramid(8, 10)
    w, o = _mk(sc)
    flip = [10, 18, 25, 27]  # boxes of rows 1-4 that touch neither the ground nor each other (a pair of two static bodies is skipped
    for s in range(50):      # by the broad phase, world.rs:79-81, and the reference then keeps its stale arbiter: DESIGN.md §8)
        if s in (15, 30):
            p = w.get_body_props()
            for i in flip:
                if s == 15:
                    p[i, 1], p[i, 3] = 0.0, 0.0          # static (kinematic: keeps its velocity, quirk Q-16)
                else:
                    p[i, 1], p[i, 3] = 0.1, 6.0          # dynamic again
            w.set_body_props(p)
            o.set_body_props(p)
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"step {s}")
    _assert_arbiters_equal(w, o)


def test_kernel_timing_api_reports_the_launch_chain():
    """sylt_world_set_kernel_timing / get_kernel_times (diagnostics used by bench.py): eight launches per step, positive durations;
    switching it off again leaves the results untouched."""
    sc = scenes.pyramid(10, 10)
    w, o = _mk(sc)
    w.set_kernel_timing(True)
    for _ in range(5):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
    kt = w.kernel_times()
    assert list(kt) == list(w.KERNEL_NAMES) and all(v > 0.0 for v in kt.values()), kt
    w.set_kernel_timing(False)
    for _ in range(5):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
    assert w.kernel_times() == {}
    _assert_states_equal(w.get_bodies(), o.get_bodies(), "kernel timing on / off")


def test_region_solver_two_ctas_per_sm(monkeypatch):
    """SYLT_REGION_CTAS_PER_SM=2 (opt-in): two half-size region CTAs per SM — half the threads, half the shared memory, half-size tables
    and bins; same results."""
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGION_CTAS_PER_SM", "2")
    monkeypatch.setenv("SYLT_REGIONS_PERIOD", "3")
    _run_lockstep(scenes.pyramid(7, 10), 40, check_every=4)
    w = _lockstep_big(scenes.box_drop(20000, seed=5, iterations=10), 150, 4)
    assert w.get_stats()["arbiters"] > 10000


def test_region_solver_prestep_keeps_the_signs_of_zero_velocities(monkeypatch):
    """With warm starting off every accumulated impulse is zero when Arbiter::pre_step applies it (arbiter.rs:216-223), so the region
    solver computes the contact constants in one flat loop instead of an ordered sweep — unless some velocity is exactly -0.0: subtracting
    a zero impulse can turn it into +0.0, and then the ordered sweep runs.  Velocities must match the oracle including the sign of zeros."""
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGIONS_COUNT", "4")
    sc = scenes.pyramid(6, 10)
    dyn = np.nonzero(sc.bodies["mass"] < scenes.F32_MAX)[0]
    sc.bodies["vx"][dyn[::2]] = -0.0
    sc.bodies["angular_velocity"][dyn[::3]] = -0.0
    sc.bodies["y"][dyn] -= 0.02  # in contact from the first step on
    w, o = _mk(sc)
    for s in range(12):
        if s == 6:  # ... and once more in the middle of the run
            st = w.get_bodies()
            st["vx"][dyn[1::2]] = -0.0
            w.set_bodies(st)
            o.set_bodies(st)
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        g, r = w.get_bodies(), o.get_bodies()
        _assert_states_equal(g, r, f"step {s}")
        for f in ("vx", "vy", "angular_velocity"):
            assert np.array_equal(np.signbit(g[f]), np.signbit(r[f])), (s, f, np.nonzero(np.signbit(g[f]) != np.signbit(r[f]))[0][:8])


def test_region_solver_survives_a_body_shot_far_away(monkeypatch):
    """One body far outside the pile must not collapse the spatial partition (k_resort trims its grid to the bulk of the
    bodies): the regions stay compact (most arbiters interior), the step succeeds and stays bit-identical to the oracle."""
    from sylt2d_b200 import World
    monkeypatch.setenv("SYLT_REGIONS", "2")
    monkeypatch.setenv("SYLT_REGIONS_PERIOD", "1")  # re-cut (and re-sort) on every step
    sc = scenes.box_drop(3000, seed=1)
    w = World(sc.gravity, sc.iterations, device=0)
    w.load_scene(sc)
    w.step_n(sc.dt, 150)
    st = w.get_bodies()
    st["x"][1234], st["y"][1234], st["vx"][1234] = 3.0e7, -2.0e7, 1.0e5
    o = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1, broadphase=1)
    o.load_scene(sc)
    o.set_bodies(st)
    w = World(sc.gravity, sc.iterations, device=0)  # fresh on both sides: no arbiters (warm-start impulses) from before
    w.load_scene(sc)
    w.set_bodies(st)
    for s in range(4):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"outlier step {s}")
    ga, _ = w.get_arbiters()
    assert ga.shape[0] > 2000 and (ga["color"] < 32).mean() > 0.6


# ------------------------------------------------------------------ error paths: exact state after Err, transparent capacity growth
def _bridge_with_static_joint(iterations=10, softness=1.0):
    """The suspension bridge (16 plank-to-ground joints) with a joint between two STATIC bodies in the middle of the joint Vec:
    with softness > 0 its K = softness * I is invertible, with softness == 0 it is the zero matrix (NoInverse, quirk Q-15)."""
    sc = scenes.bridge(iterations)
    bodies = np.concatenate([sc.bodies, sc.bodies[:1].copy()])
    bodies["id"][-1] = bodies["id"].max() + 1
    bodies["x"][-1], bodies["y"][-1] = 40.0, 30.0           # a second static body, far away from everything
    extra = np.zeros(1, dtype=scenes.JOINT_DESC)
    extra["id1"], extra["id2"] = bodies["id"][0], bodies["id"][-1]
    extra["anchor_x"], extra["anchor_y"], extra["softness"], extra["bias_factor"] = 20.0, 10.0, softness, 0.2
    joints = np.concatenate([sc.joints[:8], extra, sc.joints[8:]])
    return scenes.Scene("bridge+static-joint", bodies, joints, sc.verts, sc.gravity, sc.iterations, sc.dt, sc.ctx), 8


def _assert_joints_equal(gj, rj, what):
    for f in gj.dtype.names:
        assert np.array_equal(gj[f], rj[f]), f"{what}: joint field {f} differs: {gj[f]} vs {rj[f]}"


def test_no_inverse_mid_step_state_matches_reference_semantics():
    """World::step returns Err at the first joint (Vec order) whose K is singular (world.rs:127-129): force integration, every
    arbiter pre-step and the joints BEFORE it have happened, the joints behind it, the iterations and the integration have not.
    The state after the error — bodies, forces, every joint field, the matrix — must equal the oracle's, step after step."""
    from sylt2d_b200 import SyltError
    sc, jbad = _bridge_with_static_joint()
    w, o = _mk(sc)
    for s in range(25):                                    # warm starting is on: the joints accumulate impulses
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
    _assert_states_equal(w.get_bodies(), o.get_bodies(), "before the error")
    w.set_joint_params(jbad, 0.0, 0.2)                     # the caller writes the pub field: K becomes the zero matrix
    o.set_joint_params(jbad, 0.0, 0.2)
    w.add_force(3, (5.0, -2.0))                            # forces are NOT cleared by a failed step (world.rs:148-149 is not reached)
    o.add_force(3, (5.0, -2.0))
    for rep in range(3):
        with pytest.raises(SyltError) as ei:
            w.step(sc.dt)
        assert ei.value.code == 1
        keys, jo = w.solve_order()
        with pytest.raises(orc.OracleError) as eo:
            o.step_ordered(sc.dt, keys, jo)
        assert eo.value.code == 1
        g, r = w.get_bodies(), o.get_bodies()
        for f in g.dtype.names:                            # incl. fx, fy, torque
            assert np.array_equal(g[f], r[f]), f"after Err #{rep}: {f}"
        _assert_joints_equal(w.get_joints(), o.get_joints(), f"after Err #{rep}")
        _assert_arbiters_equal(w, o, solve_fields=False, what=f"after Err #{rep}")
        jg, mg = w.last_error_matrix()
        jr_, mr = o.last_error_matrix()
        assert jg == jr_ == jbad and np.array_equal(mg, mr)
    # the joints behind the failing one kept the impulses of the last good step; the ones before it were warm-started again
    gj = w.get_joints()
    assert np.abs(gj["px"][jbad + 1:]).max() > 0 and np.abs(gj["px"][:jbad]).max() > 0
    # step_n stops at the first Err, exactly like a caller's `for _ in 0..n { world.step(dt)?; }`
    w2, o2 = _mk(sc)
    w2.step_n(sc.dt, 25)
    w2.set_joint_params(jbad, 0.0, 0.2)
    w3, _ = _mk(sc)
    w3.step_n(sc.dt, 25)
    w3.set_joint_params(jbad, 0.0, 0.2)
    with pytest.raises(SyltError):
        w2.step_n(sc.dt, 4)
    with pytest.raises(SyltError):
        w3.step(sc.dt)
    _assert_states_equal(w2.get_bodies(), w3.get_bodies(), "step_n after Err")
    # the repaired world steps on (softness back to 1)
    w.set_joint_params(jbad, 1.0, 0.2)
    o.set_joint_params(jbad, 1.0, 0.2)
    for s in range(5):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"after the repair, step {s}")


def test_no_inverse_in_a_batch_world_matches_oracle():
    from sylt2d_b200 import Batch
    sc, jbad = _bridge_with_static_joint(softness=0.0)     # singular from the first step
    bodies, boff, joints, joff = scenes.tile_worlds(sc, 3, seed=1, omega_jitter=0.5)
    b = Batch(bodies, boff, joints, joff, sc.gravity, sc.iterations, device=0, ctx=sc.ctx)
    b.step_n(sc.dt, 2)
    assert (b.get_errors() == 1).all()
    nb = sc.num_bodies
    for wi in range(3):
        o = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1)
        o.load_scene(scenes.Scene(sc.name, bodies[wi * nb:(wi + 1) * nb], sc.joints, sc.verts, sc.gravity, sc.iterations, sc.dt, sc.ctx))
        with pytest.raises(orc.OracleError):
            o.step_ordered(sc.dt, np.zeros((0, 2), dtype=np.uint64), b.get_joint_order(wi))
        _assert_states_equal(b.get_bodies(wi, 1), o.get_bodies(), f"batch world {wi} after Err")
        gj, rj = b.get_joints(wi), o.get_joints()
        for f in ("r1x", "r1y", "r2x", "r2y", "px", "py", "bias_x", "bias_y"):
            assert np.array_equal(gj[f], rj[f]), (wi, f)
        assert np.abs(gj["r1x"][:jbad + 1]).max() > 0 and not gj["r1x"][jbad + 1:].any()  # joints behind the failing one: untouched


def test_capacity_overflow_is_transparent():
    """The reference's arbiter container is a HashMap (no limit).  Device buffers that turn out too small are grown and the step
    is repeated; a failed attempt changes nothing, so the result is bit-identical to a world that never overflowed."""
    from sylt2d_b200 import World
    sc = scenes.pyramid(10, 10)
    w, o = _mk(sc)
    w.set_capacity(24, 12, 20)                             # far too small for 56 bodies
    for s in range(70):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"tiny capacity, step {s}")
    _assert_arbiters_equal(w, o, what="tiny capacity")
    assert w.get_stats()["arbiters"] > 60
    # multi-step calls: the overflow happens in the middle of a step_n; the steps behind it are repeated too
    ref = World(sc.gravity, sc.iterations)
    ref.load_scene(sc)
    ref.step_n(sc.dt, 70)
    for chunks in ((70,), (20, 50), (5, 5, 60)):
        v = World(sc.gravity, sc.iterations)
        v.load_scene(sc)
        v.set_capacity(24, 12, 20)
        for n in chunks:
            v.step_n(sc.dt, n)
        _assert_states_equal(v.get_bodies(), ref.get_bodies(), f"tiny capacity, step_n chunks {chunks}")
    # a bigger, region-solved world with polygons
    sc = scenes.mixed_pile(1500, seed=3)
    a, bb = World(sc.gravity, sc.iterations), World(sc.gravity, sc.iterations)
    a.load_scene(sc)
    bb.load_scene(sc)
    bb.set_capacity(2000, 600, 900)
    a.step_n(sc.dt, 90)
    bb.step_n(sc.dt, 90)
    _assert_states_equal(a.get_bodies(), bb.get_bodies(), "mixed pile, grown capacity")
    assert a.get_stats()["arbiters"] == bb.get_stats()["arbiters"] > 600


def test_pub_body_and_joint_fields_written_between_steps():
    """Body.friction / mass / inv_mass / moi / inv_moi / width and Joint.local_anchor_* are pub fields the reference reads on every
    step (world.rs:115-118, arbiter.rs:128,198-211, collide.rs:141-142, joint.rs:67-68): a caller may overwrite them between steps."""
    sc = scenes.demo8(10) if False else scenes.pyramid(6, 10)
    w, o = _mk(sc)
    for s in range(60):
        if s == 20:
            p = w.get_body_props()
            assert np.array_equal(p, o.get_body_props())
            p[3, 6] = 0.9                                   # friction: only arbiters created from now on see it (quirk Q-17)
            p[5, 0], p[5, 1], p[5, 2], p[5, 3] = 40.0, 1.0 / 40.0, 40.0 * 2.0 / 12.0, 12.0 / (40.0 * 2.0)   # a heavier box
            p[7, 4], p[7, 5] = 1.2, 0.8                     # width: collide() reads it every step
            p[9, 1], p[9, 3] = 0.0, 0.0                     # inv_mass = inv_moi = 0: the body becomes static (kinematic: it keeps moving, quirk Q-16)
            w.set_body_props(p)
            o.set_body_props(p)
        if s == 40:
            p = w.get_body_props()
            p[9, 1], p[9, 3] = 0.1, 6.0                     # ... and dynamic again
            w.set_body_props(p[9:10], first=9)
            o.set_body_props(p[9:10], first=9)
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"props step {s}")
    _assert_arbiters_equal(w, o, what="props")
    assert np.array_equal(w.get_body_props(), o.get_body_props())
    # joints: local anchors edited after Joint::new
    sc = scenes.multi_pendulum(10)
    w, o = _mk(sc)
    for s in range(50):
        if s == 15:
            gj = w.get_joints()
            la1, la2 = (float(gj["la1x"][3]) + 0.1, float(gj["la1y"][3])), (float(gj["la2x"][3]), float(gj["la2y"][3]) - 0.05)
            w.set_joint_local_anchors(3, la1, la2)
            o.set_joint_local_anchors(3, la1, la2)
            w.set_joint_params(5, 0.02, 0.1)
            o.set_joint_params(5, 0.02, 0.1)
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"joint fields step {s}")
    _assert_joints_equal(w.get_joints(), o.get_joints(), "joint fields")


def test_rust_wrapper_call_sequence_replayed_in_cpp(tmp_path):
    """rust/src/world.rs cannot be compiled here (no rustc): tests/cpp/wrapper_sequence.cpp replays its per-step protocol in C++ —
    host-side bodies / joints with pub fields, diffed against a mirror (upload_dirty), sylt_world_step, mirror_back, lazy arbiter
    fetch — incl. a bomb launch between steps, mutated state / properties / joint fields and a context toggle, and compares every
    step with the oracle bit for bit."""
    import subprocess
    import __graft_entry__ as g
    from tests.test_abi import ROOT
    g.build()
    exe = os.path.join(str(tmp_path), "wrapper_sequence")
    lib, orcdir = os.path.join(ROOT, "sylt2d_b200"), os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-O1", os.path.join(ROOT, "tests", "cpp", "wrapper_sequence.cpp"), "-L", lib, "-lsylt2d_b200",
                    "-L", orcdir, "-lorc", f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{orcdir}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bit-identical" in r.stdout, r.stdout + r.stderr


def test_bodies_at_huge_coordinates_stay_exact_and_cheap():
    """The conservative AABB margin grows with |x|; at |x| ~ 1e7 a body's search window is hundreds of cells wide.  Such bodies are
    searched by walking the cell-ordered entries instead of the cells (k_rows_count's wide path): the pair set stays a superset of
    the all-pairs oracle's hits (bit-exact arbiters and states) and a step stays cheap (an exploding polygon pile used to take
    150 ms per step)."""
    import time
    b = scenes._Builder()
    scenes._ground(b)
    for r in range(6):
        for c in range(6 - r):
            b.box(1.0, 1.0, 10.0, 0.2, (-3.0 + 0.5625 * r + 1.125 * c, 0.75 + 1.05 * r))
    for k in range(60):                                     # far away: ulp(1e7) = 1, so neighbours one unit apart overlap by construction
        b.box(1.5, 1.5, 5.0, 0.3, (1.0e7 + (k // 2) * 6.0 + (k % 2) * 1.0, 3.0e6 + (k % 3)), rot=0.1 * k, vel=(0.5 * (k % 5), -1.0))
    b.box(2.0, 2.0, 5.0, 0.3, (-4.0e8, 2.0e8), vel=(1.0e4, 0.0))
    sc = b.scene("far_bodies", iterations=10)
    w, o = _mk(sc)                                          # oracle: the reference's all-pairs loop
    t0 = time.perf_counter()
    for s in range(6):
        w.step(sc.dt)
        keys, jo = w.solve_order()
        o.step_ordered(sc.dt, keys, jo)
        _assert_states_equal(w.get_bodies(), o.get_bodies(), f"far bodies step {s}")
    ga = _assert_arbiters_equal(w, o, what="far bodies")
    assert (ga["body1"] > 21).sum() >= 20                   # the far bodies do touch one another
    assert time.perf_counter() - t0 < 5.0
