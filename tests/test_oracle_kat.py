"""Pins the CPU oracle against every known-answer the reference's own unit tests hold for this path
(SURVEY.md §4 / §8c), and cross-validates it against the independent numpy-f32 restatement.

Reference tests restated here:
  src/collide.rs:306-593      four box pairs — contact COUNT only (== 0 / > 0)
  src/math_utils.rs:256-311   Vec2 / Cross / Mat2x2 spot values
  src/errors.rs:42-66         singular matrix -> NoInverse carrying the matrix
  src/body.rs:308-313         add_force accumulates
The "survey cross-check" table is an extra regression net, not reference ground truth.
"""
import math

import numpy as np
import pytest

from oracle import orc, narrow_np

F = np.float32
RAD45 = float(F(45.0) * (F(np.pi) / F(180.0)))  # 45.0_f32.to_radians()


def box(w, pos, rot=0.0):
    return dict(shape=0, width=w, pos=pos, rot=rot)


def poly(verts, pos, rot=0.0):
    return dict(shape=1, verts=np.asarray(verts, dtype=np.float32), pos=pos, rot=rot)


REF_COUNT_TESTS = [  # (A, B, predicate) — src/collide.rs test_no_overlap / test_overlap / _at_angle / _and_fixed
    (box((1, 1), (10, 1)), box((1, 1), (15, 5)), lambda n: n == 0),
    (box((2, 2), (11, 3)), box((2, 2), (12, 2)), lambda n: n > 0),
    (box((4, 4), (12, 0), RAD45), box((4, 4), (15.5, 1), -RAD45), lambda n: n > 0),
    (box((4, 4), (14, 2), RAD45), box((4, 4), (18, 2)), lambda n: n > 0),
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", range(4))
def test_reference_contact_counts(case, mode):
    a, b, pred = REF_COUNT_TESTS[case]
    assert pred(len(orc.collide_pair(a, b, sincos_mode=mode)))


SURVEY_TABLE = [  # SURVEY.md §4 cross-check vectors: (A, B, [(sep, normal, pos, edges)])
    (box((1, 1), (10, 1)), box((1, 1), (15, 5)), []),
    (box((2, 2), (11, 3)), box((2, 2), (12, 2)), [(-1, (1, 0), (12, 3), (0, 0, 1, 2)), (-1, (1, 0), (12, 2), (0, 3, 2, 0))]),
    (box((4, 4), (12, 0), RAD45), box((4, 4), (15.5, 1), -RAD45),
     [(-0.81801987, (0.70710677, 0.70710677), (13.25, 1.5784273), (0, 0, 2, 3)),
      (-0.81801987, (0.70710677, 0.70710677), (14.828427, 2.98e-7), (0, 3, 3, 0))]),
    (box((4, 4), (14, 2), RAD45), box((4, 4), (18, 2)), [(-0.82842636, (1, 0), (16, 2), (3, 4, 0, 0))]),
    (box((3, 2), (8.5, -0.5), 0.2), box((3, 2), (10, 1)),
     [(-0.77807057, (-0.0, 1), (9.771431, 0), (4, 1, 0, 0)), (-0.52033883, (-0.0, 1), (8.5, 0), (1, 0, 0, 2))]),
    (box((100, 20), (0, -10)), box((1, 1), (0, 0)), [(-0.5, (-0.0, 1), (-0.5, 0), (0, 0, 2, 3)), (-0.5, (-0.0, 1), (0.5, 0), (0, 0, 3, 4))]),
    (box((4, 4), (12, 0), RAD45), box((2, 2), (15.5, 1), RAD45), []),
    (box((4, 4), (14, 2), RAD45), box((2, 2), (18, 2), RAD45),
     [(-0.17157269, (0.7071, 0.7071), (16.707108, 2.1213202), (0, 0, 1, 2)), (-0.17157173, (0.7071, 0.7071), (16.828426, 1.9999993), (0, 3, 2, 0))]),
    (box((2, 2), (1, 1), RAD45), box((2, 2), (5, 1), RAD45), []),
]


@pytest.mark.parametrize("case", range(len(SURVEY_TABLE)))
def test_survey_crosscheck_table(case):
    a, b, exp = SURVEY_TABLE[case]
    got = orc.collide_pair(a, b, sincos_mode=0)
    assert len(got) == len(exp)
    for g, (sep, n, p, e) in zip(got, exp):
        assert abs(g["separation"] - sep) < 2e-6
        assert abs(g["nx"] - n[0]) < 1e-4 and abs(g["ny"] - n[1]) < 1e-4
        assert abs(g["px"] - p[0]) < 2e-6 * max(1, abs(p[0])) and abs(g["py"] - p[1]) < 2e-6 * max(1, abs(p[1]))
        assert (g["in_edge_1"], g["out_edge_1"], g["in_edge_2"], g["out_edge_2"]) == e
        assert g["feature_value"] == 0  # quirk Q-2: never written


def test_math_spot_values():
    # math_utils.rs:258-263 add_sub_ops
    p1, p2 = np.array([0.0, 1.0], F), np.array([3.2, 2.9], F)
    assert np.all(np.abs((p1 + p2) - np.array([3.2, 3.9], F)) < np.finfo(F).eps)
    assert np.all(np.abs((p1 - p2) - np.array([-3.2, -1.9000001], F)) < np.finfo(F).eps)
    # math_utils.rs:286-293 test_mat: invert() of the 90 degree rotation (det == 1 hides quirk Q-3)
    s, c = F(0), F(0)
    import ctypes as C
    for mode in (0, 1):
        ss, cc = C.c_float(), C.c_float()
        orc.lib().orc_sincos(mode, float(F(np.pi) / F(2.0)), C.byref(ss), C.byref(cc))
        s, c = F(ss.value), F(cc.value)
        m = np.array([c, -s, s, c], F)  # (col1.x, col2.x, col1.y, col2.y)
        e, inv = orc.mat_invert(m)
        assert e == 0 and inv[1] == 1.0          # invert().col2.x == 1.0
        assert m[1] == -1.0                      # transpose().col1.y == col2.x == -1.0
        # math_utils.rs:296-310 test_mat_ops
        orc.lib().orc_sincos(mode, float(F(np.pi) / F(4.0)), C.byref(ss), C.byref(cc))
        s1, c1 = F(ss.value), F(cc.value)
        orc.lib().orc_sincos(mode, float(-(F(np.pi) / F(4.0))), C.byref(ss), C.byref(cc))
        s2, c2 = F(ss.value), F(cc.value)
        assert F(c1 * F(-s2)) + F(F(-s1) * c2) == 0.0            # (mat1*mat2).col2.x == 0
        assert F(c1 + c2) == F(math.sqrt(2.0))                    # (mat1+mat2).col1.x == SQRT_2
        assert (F(c1 * F(2.0)) + F(F(-s1) * F(3.0))) - F(1 / math.sqrt(2.0)) < np.finfo(F).eps


def test_invert_quirk_q3_and_no_inverse():
    e, _ = orc.mat_invert([1.0, 2.0, 2.0, 4.0])  # errors.rs:45-48 (col1=(1,2), col2=(2,4))
    assert e == 1  # SYLT_ERR_NO_INVERSE
    e, inv = orc.mat_invert([2.0, 0.0, 0.0, 4.0])  # det = 8: adjugate * det, not / det
    assert e == 0 and list(inv) == [32.0, -0.0, -0.0, 16.0]


def test_no_inverse_mid_step_carries_matrix():
    # joint.rs:95 + world.rs:127-129 (quirk Q-15): two static bodies -> K == 0 -> NoInverse
    w = orc.OracleWorld(iterations=10)
    fm = float(np.finfo(F).max)
    w.add_box(1, (1, 1), fm)
    w.add_box(2, (1, 1), fm, pos=(3, 0))
    w.add_joint(1, 2, (1.5, 0.0))
    with pytest.raises(orc.OracleError) as ei:
        w.step(1 / 60)
    assert ei.value.code == 1
    j, m = w.last_error_matrix()
    assert j == 0 and list(m) == [0, 0, 0, 0]


def test_add_force_accumulates():
    w = orc.OracleWorld()
    w.add_box(1, (1, 1), 1.0)
    w.add_force(0, (2.0, 5.3))
    w.add_force(0, (1.0, 0.0))
    b = w.get_bodies()[0]
    assert b["fx"] == F(3.0) and b["fy"] == F(5.3)


def test_body_order_error_q6():
    w = orc.OracleWorld()
    w.add_box(5, (1, 1), 1.0)
    w.add_box(2, (1, 1), 1.0, pos=(10, 0))
    with pytest.raises(orc.OracleError) as ei:
        w.step(1 / 60)
    assert ei.value.code == 3


# ------------------------------------------------------------------ independent restatement cross-check
def _sincos(mode):
    import ctypes as C
    L = orc.lib()

    def f(a):
        s, c = C.c_float(), C.c_float()
        L.orc_sincos(mode, float(a), C.byref(s), C.byref(c))
        return s.value, c.value
    return f


def _same_contacts(got, exp):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        for k, v in (("separation", e["sep"]), ("nx", e["normal"][0]), ("ny", e["normal"][1]), ("px", e["pos"][0]), ("py", e["pos"][1])):
            assert np.float32(g[k]).tobytes() == np.float32(v).tobytes() or (g[k] == v), (k, g[k], v)
        assert (g["in_edge_1"], g["out_edge_1"], g["in_edge_2"], g["out_edge_2"]) == tuple(e["edges"])


@pytest.mark.parametrize("mode", [0, 1])
def test_box_narrow_phase_bit_exact_vs_numpy(mode):
    rng = np.random.default_rng(7)
    sc = _sincos(mode)
    hits = 0
    for _ in range(600):
        wa, wb = rng.uniform(0.3, 3.0, 2), rng.uniform(0.3, 3.0, 2)
        pa = rng.uniform(-3, 3, 2)
        pb = pa + rng.uniform(-2.5, 2.5, 2)
        ra, rb = rng.uniform(-7, 7, 2)
        a, b = box(tuple(F(wa)), tuple(F(pa)), float(F(ra))), box(tuple(F(wb)), tuple(F(pb)), float(F(rb)))
        got = orc.collide_pair(a, b, sincos_mode=mode)
        exp = narrow_np.collide_boxes(a["width"], a["pos"], a["rot"], b["width"], b["pos"], b["rot"], sc)
        _same_contacts(got, exp)
        hits += len(got) > 0
    assert hits > 100


def _rand_poly(rng):
    n = int(rng.integers(3, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    r = rng.uniform(0.4, 1.2)
    return np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)


@pytest.mark.parametrize("mode", [0, 1])
def test_polygon_narrow_phase_bit_exact_vs_numpy(mode):
    rng = np.random.default_rng(11)
    sc = _sincos(mode)
    hits = 0
    for k in range(250):
        va = _rand_poly(rng)
        pa = F(rng.uniform(-3, 3, 2))
        pb = F(pa + rng.uniform(-1.5, 1.5, 2))
        ra, rb = float(F(rng.uniform(-7, 7))), float(F(rng.uniform(-7, 7)))
        a = poly(va, tuple(pa), ra)
        if k % 3 == 0:  # Box-Polygon pairs take the polygon path with the box's 4 stored vertices (quirk Q-11)
            wb = tuple(F(rng.uniform(0.4, 2.0, 2)))
            b = box(wb, tuple(pb), rb)
            vb = narrow_np.BOX_VERTS(wb)
        else:
            vb = _rand_poly(rng)
            b = poly(vb, tuple(pb), rb)
        got = orc.collide_pair(a, b, sincos_mode=mode)
        exp = narrow_np.collide_polys(va, pa, ra, vb, pb, rb, sc)
        _same_contacts(got, exp)
        hits += len(got) > 0
    assert hits > 50


def test_collision_debug_polygon_setups():
    # examples/collision-debug/src/main.rs:145-193: hexagon vs pentagon at the origin, hexagon vs rotated box
    from sylt2d_b200.scenes import HEXAGON, PENTAGON
    sc = _sincos(0)
    got = orc.collide_pair(poly(HEXAGON, (0, 0)), poly(PENTAGON, (0, 0)), sincos_mode=0)
    exp = narrow_np.collide_polys(HEXAGON, (0, 0), 0.0, PENTAGON, (0, 0), 0.0, sc)
    assert len(got) >= 3
    _same_contacts(got, exp)
    b = box((2.0, 2.0), (1.0, 0.5), 0.3)
    got = orc.collide_pair(poly(HEXAGON, (0, 0)), b, sincos_mode=0)
    exp = narrow_np.collide_polys(HEXAGON, (0, 0), 0.0, narrow_np.BOX_VERTS((2.0, 2.0)), (1.0, 0.5), 0.3, sc)
    assert len(got) >= 3
    _same_contacts(got, exp)


# ------------------------------------------------------------------ oracle self-consistency at scale
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_grid_broad_phase_equals_all_pairs(seed):
    """The oracle's grid candidate mode (used for full-size parity runs and the algorithm-matched CPU number) must
    produce exactly the arbiter set of the reference's all-pairs loop (src/world.rs:74-103)."""
    from sylt2d_b200 import scenes
    sc = scenes.mixed_pile(600, seed=seed, across=30)
    a = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=0, broadphase=0)
    b = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=0, broadphase=1)
    a.load_scene(sc)
    b.load_scene(sc)
    for s in range(110):
        a.step(sc.dt)
        b.step(sc.dt)
    assert a.get_bodies().tobytes() == b.get_bodies().tobytes()
    (aa, ac), (ba, bc) = a.get_arbiters(), b.get_arbiters()
    assert aa.shape[0] > 100 and aa.tobytes() == ba.tobytes() and ac.tobytes() == bc.tobytes()


# ------------------------------------------------------------------ World::step: C++ oracle vs the numpy-f32 restatement
def _np_vs_cpp(scene, steps, mode=1, ctx=None):
    """Both restatements step the same scene in canonical (id1, id2) order; every body state and every contact's
    accumulated impulses must agree bit for bit after every step."""
    from oracle import step_np
    if ctx is not None:
        scene.ctx = ctx
    o = orc.OracleWorld(scene.gravity, scene.iterations, sincos_mode=mode)
    o.load_scene(scene)
    w = step_np.NpWorld(scene.gravity, scene.iterations, _sincos(mode), scene.ctx)
    w.load_scene(scene, o.get_body_props())
    n_contacts = 0
    for s in range(steps):
        o.step(scene.dt)
        assert w.step(scene.dt)
        got = o.get_bodies()
        exp = w.state()
        for k, f in enumerate(("x", "y", "rotation", "vx", "vy", "angular_velocity")):
            assert np.array_equal(got[f].view(np.uint32), exp[:, k].view(np.uint32)), (scene.name, s, f)
        arb, con = o.get_arbiters()
        keys = sorted(w.arbiters)
        assert [tuple(k) for k in zip(arb["id1"].tolist(), arb["id2"].tolist())] == keys, (scene.name, s)
        flat = [c for k in keys for c in w.arbiters[k]["contacts"]]
        assert len(flat) == con.shape[0]
        for f in ("pn", "pt", "mass_normal", "mass_tangent", "bias", "separation"):
            src = {"mass_normal": "mass_n", "mass_tangent": "mass_t", "separation": "sep"}.get(f, f)
            exp_f = np.array([c[src] for c in flat], dtype=np.float32)
            assert np.array_equal(con[f].view(np.uint32), exp_f.view(np.uint32)), (scene.name, s, f)
        n_contacts = max(n_contacts, len(flat))
    return n_contacts


@pytest.mark.parametrize("mode", [0, 1])
def test_step_pyramid_bit_exact_vs_numpy(mode):
    from sylt2d_b200 import scenes
    assert _np_vs_cpp(scenes.pyramid(4, 10), 70, mode) >= 16


@pytest.mark.parametrize("ctx", [(True, False, True), (False, True, True), (True, True, False), (False, False, False)])
def test_step_context_flags_bit_exact_vs_numpy(ctx):
    from sylt2d_b200 import scenes
    assert _np_vs_cpp(scenes.pyramid(3, 6), 60, 1, ctx) >= 8


def test_step_joints_bit_exact_vs_numpy():
    from sylt2d_b200 import scenes
    _np_vs_cpp(scenes.bridge(10, omega0=0.5), 40)
    _np_vs_cpp(scenes.multi_pendulum(10, omega0=0.7), 40)
    _np_vs_cpp(scenes.demo2(10), 40)  # single pendulum, hard joint


def test_step_polygons_and_boxes_bit_exact_vs_numpy():
    from sylt2d_b200 import scenes
    b = scenes._Builder()
    scenes._ground(b)
    b.polygon(scenes.HEXAGON, 5.0, 0.3, (-2.0, 1.2), 0.1)
    b.box(1.0, 1.0, 10.0, 0.2, (0.0, 0.6), 0.0)
    b.polygon(scenes.PENTAGON, 4.0, 0.4, (0.2, 2.3), -0.3, (0.0, -1.0), 0.5)
    b.box(1.5, 0.5, 6.0, 0.5, (2.0, 0.4), 0.05)
    b.polygon(scenes.regular_ngon(3, 0.8), 3.0, 0.2, (2.1, 1.6), 0.4)
    assert _np_vs_cpp(b.scene("mixed", iterations=10), 60) >= 4
