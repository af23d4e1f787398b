"""Second, independent restatement of the reference's World::step in numpy-float32 scalars.

TEST INFRASTRUCTURE, NOT PRODUCT.  Purpose (SURVEY.md §8c "Mitigation" (2)): the reference's own tests do not pin `step`,
`broad_phase`, `Arbiter::update / pre_step / apply_impulse` or the joints, and the Rust crate cannot be built here, so the
C++ oracle (oracle/sylt_oracle.cpp) is cross-validated bit-for-bit against this separately written version
(tests/test_oracle_kat.py).  It was written from the Rust sources, not from the C++ oracle.  Pure-Python loops: small
scenes only.

Follows /root/reference/src/world.rs:73-152 (broad_phase, step), src/arbiter.rs:85-298 (ArbiterKey, Arbiter::new / update /
pre_step / apply_impulse), src/joint.rs:26-130, src/body.rs:204-245 (box mass properties), src/math_utils.rs:81-106,185-199.
The narrow phases come from oracle/narrow_np.py (the same second restatement).  Polygon mass properties (Body::new_polygon,
body.rs:246-284: host-side, once per body) are taken as inputs.

Iteration order of `World::arbiters` (a HashMap in the reference, quirk Q-1): ascending (id1, id2), the C++ oracle's canonical
order.  sin/cos are injected so both restatements share one libm.
"""
import numpy as np

from . import narrow_np as nn

F = np.float32
ZERO = F(0.0)
F32_MAX = F(np.finfo(np.float32).max)


def _v(x, y):
    return (F(x), F(y))


def _cross_vv(a, b):  # Vec2 x Vec2 (math_utils.rs:81-86)
    return F(F(a[0] * b[1]) - F(a[1] * b[0]))


def _cross_vs(a, s):  # Vec2 x f32 (math_utils.rs:88-96)
    return (F(a[1] * s), F(a[0] * F(-s)))


def _cross_sv(s, a):  # f32 x Vec2 (math_utils.rs:98-106)
    return (F(F(-s) * a[1]), F(s * a[0]))


def _clamp(x, lo, hi):  # f32::clamp
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


class NpBody:
    def __init__(self, bid, shape, width, verts, mass, friction, pos, rot, vel, omega, inv_mass=None, inv_moi=None):
        self.id, self.shape = int(bid), int(shape)
        self.width = _v(*width)
        self.verts = None if verts is None else [(F(p[0]), F(p[1])) for p in verts]
        self.friction = F(friction)
        mass = F(mass)
        if inv_mass is None:  # Body::new (body.rs:204-217)
            if mass < F32_MAX:
                inv_mass = F(F(1.0) / mass)
                moi = F(F(mass * F(F(self.width[0] * self.width[0]) + F(self.width[1] * self.width[1]))) / F(12.0))
                inv_moi = F(F(1.0) / moi)
            else:
                inv_mass, inv_moi = ZERO, ZERO
        self.inv_mass, self.inv_moi = F(inv_mass), F(inv_moi)
        self.pos, self.rot = _v(*pos), F(rot)
        self.vel, self.omega = _v(*vel), F(omega)
        self.force, self.torque = _v(0.0, 0.0), ZERO

    def local_verts(self):  # Body::new stores the box corners (body.rs:218-225); polygons keep the vertices as given
        return nn.BOX_VERTS(self.width) if self.shape == 0 else self.verts


class NpWorld:
    def __init__(self, gravity, iterations, sincos, ctx=(True, False, True)):
        self.g = _v(*gravity)
        self.iterations = int(iterations)
        self.sincos = sincos
        self.accumulate, self.warm, self.poscorr = ctx  # WorldContext (world.rs:11-16; defaults :37-41)
        self.bodies, self.joints = [], []
        self.arbiters = {}  # (id1, id2) -> dict(b1, b2, friction, contacts)

    # ---- scene loading (sylt2d_b200.scenes.Scene); props = rows of (mass, inv_mass, moi, inv_moi, ...) for polygons
    def load_scene(self, scene, props=None):
        self.iterations = int(scene.iterations)
        self.accumulate, self.warm, self.poscorr = scene.ctx
        for k, b in enumerate(scene.bodies):
            verts, im, ii = None, None, None
            if b["shape"] != 0:
                verts = scene.verts[b["vert_offset"]: b["vert_offset"] + b["vert_count"]]
                im, ii = props[k][1], props[k][3]
            self.bodies.append(NpBody(b["id"], b["shape"], (b["width_x"], b["width_y"]), verts, b["mass"], b["friction"],
                                      (b["x"], b["y"]), b["rotation"], (b["vx"], b["vy"]), b["angular_velocity"], im, ii))
        for j in scene.joints:
            self.add_joint(int(j["id1"]), int(j["id2"]), (j["anchor_x"], j["anchor_y"]), j["softness"], j["bias_factor"])

    def _find(self, bid):
        for k, b in enumerate(self.bodies):
            if b.id == bid:
                return k
        raise KeyError(bid)

    def add_joint(self, id1, id2, anchor, softness=0.0, bias_factor=0.2):  # Joint::new (joint.rs:26-55)
        b1, b2 = self.bodies[self._find(id1)], self.bodies[self._find(id2)]
        anchor = _v(*anchor)
        rt1, rt2 = nn._T(nn._rot(b1.rot, self.sincos)), nn._T(nn._rot(b2.rot, self.sincos))
        self.joints.append({"b1": b1, "b2": b2, "la1": nn._mul_mv(rt1, nn._sub(anchor, b1.pos)), "la2": nn._mul_mv(rt2, nn._sub(anchor, b2.pos)),
                            "softness": F(softness), "bias_factor": F(bias_factor), "p": _v(0, 0), "bias": _v(0, 0),
                            "r1": _v(0, 0), "r2": _v(0, 0), "m": (ZERO, ZERO, ZERO, ZERO)})

    # ---- Arbiter::new (arbiter.rs:117-136) for bodies already in id order
    def _collide(self, a, b):
        if a.shape == 0 and b.shape == 0:
            cs = nn.collide_boxes(a.width, a.pos, a.rot, b.width, b.pos, b.rot, self.sincos)
        else:
            cs = nn.collide_polys(a.local_verts(), a.pos, a.rot, b.local_verts(), b.pos, b.rot, self.sincos)
        return [{"pos": c["pos"], "normal": c["normal"], "sep": c["sep"], "pn": ZERO, "pt": ZERO, "pnb": ZERO,
                 "mass_n": ZERO, "mass_t": ZERO, "bias": ZERO} for c in cs]

    def broad_phase(self):  # world.rs:73-105
        n = len(self.bodies)
        for i in range(n):
            bi = self.bodies[i]
            for j in range(i + 1, n):
                bj = self.bodies[j]
                if bi.inv_mass == ZERO and bj.inv_mass == ZERO:
                    continue
                assert bi.id < bj.id, "bodies must be added in increasing id order (arbiter.rs:120-122 swaps them otherwise)"
                contacts = self._collide(bi, bj)
                key = (bi.id, bj.id)
                if contacts:
                    old = self.arbiters.get(key)
                    if old is None:
                        with np.errstate(all="ignore"):
                            fr = F(np.sqrt(F(bi.friction * bj.friction)))
                        self.arbiters[key] = {"b1": bi, "b2": bj, "friction": fr, "contacts": contacts}
                    else:  # Arbiter::update (arbiter.rs:137-181): FeaturePair.value is never written, so every new
                        #    contact matches old contact 0
                        c_old = old["contacts"][0]
                        for c in contacts:
                            if self.warm:
                                c["pn"], c["pt"], c["pnb"] = c_old["pn"], c_old["pt"], c_old["pnb"]
                        old["contacts"] = contacts
                else:
                    self.arbiters.pop(key, None)

    def _pre_step_arbiter(self, arb, inv_dt):  # arbiter.rs:182-228
        k_allowed = F(0.01)
        k_bias = F(0.2) if self.poscorr else ZERO
        b1, b2 = arb["b1"], arb["b2"]
        for c in arb["contacts"]:
            n = c["normal"]
            r1, r2 = nn._sub(c["pos"], b1.pos), nn._sub(c["pos"], b2.pos)
            rn1, rn2 = nn._dot(r1, n), nn._dot(r2, n)
            k_normal = F(b1.inv_mass + b2.inv_mass)
            k_normal = F(k_normal + F(F(b1.inv_moi * F(nn._dot(r1, r1) - F(rn1 * rn1))) + F(b2.inv_moi * F(nn._dot(r2, r2) - F(rn2 * rn2)))))
            c["mass_n"] = F(F(1.0) / k_normal)
            t = _cross_vs(n, F(1.0))
            rt1, rt2 = nn._dot(r1, t), nn._dot(r2, t)
            k_tangent = F(b1.inv_mass + b2.inv_mass)
            k_tangent = F(k_tangent + F(F(b1.inv_moi * F(nn._dot(r1, r1) - F(rt1 * rt1))) + F(b2.inv_moi * F(nn._dot(r2, r2) - F(rt2 * rt2)))))
            c["mass_t"] = F(F(1.0) / k_tangent)
            c["bias"] = F(F(F(-k_bias) * inv_dt) * F(min(ZERO, F(c["sep"] + k_allowed))))
            if self.accumulate:
                p = nn._add(nn._scale(n, c["pn"]), nn._scale(t, c["pt"]))
                b1.vel = nn._sub(b1.vel, nn._scale(p, b1.inv_mass))
                b1.omega = F(b1.omega - F(b1.inv_moi * _cross_vv(r1, p)))
                b2.vel = nn._add(b2.vel, nn._scale(p, b2.inv_mass))
                b2.omega = F(b2.omega + F(b2.inv_moi * _cross_vv(r2, p)))

    @staticmethod
    def _dv(b1, b2, r1, r2):  # v2 + w2 x r2 - v1 - w1 x r1, left to right (arbiter.rs:240-242)
        return nn._sub(nn._sub(nn._add(b2.vel, _cross_sv(b2.omega, r2)), b1.vel), _cross_sv(b1.omega, r1))

    def _apply_arbiter(self, arb):  # arbiter.rs:229-298
        b1, b2 = arb["b1"], arb["b2"]
        for c in arb["contacts"]:
            n = c["normal"]
            r1, r2 = nn._sub(c["pos"], b1.pos), nn._sub(c["pos"], b2.pos)
            vn = nn._dot(self._dv(b1, b2, r1, r2), n)
            d_pn = F(c["mass_n"] * F(F(-vn) + c["bias"]))
            if self.accumulate:
                pn0 = c["pn"]
                c["pn"] = F(max(F(pn0 + d_pn), ZERO))
                d_pn = F(c["pn"] - pn0)
            else:
                d_pn = F(max(ZERO, d_pn))
            pn = nn._scale(n, d_pn)
            b1.vel = nn._sub(b1.vel, nn._scale(pn, b1.inv_mass))
            b1.omega = F(b1.omega - F(b1.inv_moi * _cross_vv(r1, pn)))
            b2.vel = nn._add(b2.vel, nn._scale(pn, b2.inv_mass))
            b2.omega = F(b2.omega + F(b2.inv_moi * _cross_vv(r2, pn)))
            t = _cross_vs(n, F(1.0))
            vt = nn._dot(self._dv(b1, b2, r1, r2), t)
            d_pt = F(c["mass_t"] * F(-vt))
            if self.accumulate:
                max_pt = F(arb["friction"] * c["pn"])
                old = c["pt"]
                c["pt"] = _clamp(F(old + d_pt), F(-max_pt), max_pt)
                d_pt = F(c["pt"] - old)
            else:
                max_pt = F(arb["friction"] * d_pn)
                d_pt = _clamp(d_pt, F(-max_pt), max_pt)
            pt = nn._scale(t, d_pt)
            b1.vel = nn._sub(b1.vel, nn._scale(pt, b1.inv_mass))
            b1.omega = F(b1.omega - F(b1.inv_moi * _cross_vv(r1, pt)))
            b2.vel = nn._add(b2.vel, nn._scale(pt, b2.inv_mass))
            b2.omega = F(b2.omega + F(b2.inv_moi * _cross_vv(r2, pt)))

    def _pre_step_joint(self, j, inv_dt):  # joint.rs:57-115; returns False on a singular K (math_utils.rs:191-192)
        b1, b2 = j["b1"], j["b2"]
        r1 = nn._mul_mv(nn._rot(b1.rot, self.sincos), j["la1"])
        r2 = nn._mul_mv(nn._rot(b2.rot, self.sincos), j["la2"])
        j["r1"], j["r2"] = r1, r2
        s = F(b1.inv_mass + b2.inv_mass)
        k1 = (s, ZERO, ZERO, s)  # (col1.x, col1.y, col2.x, col2.y)
        i1, i2 = b1.inv_moi, b2.inv_moi
        k2 = (F(F(i1 * r1[1]) * r1[1]), F(F(F(-i1) * r1[0]) * r1[1]), F(F(F(-i1) * r1[0]) * r1[1]), F(F(i1 * r1[0]) * r1[0]))
        k3 = (F(F(i2 * r2[1]) * r2[1]), F(F(F(-i2) * r2[0]) * r2[1]), F(F(F(-i2) * r2[0]) * r2[1]), F(F(i2 * r2[0]) * r2[0]))
        k = [F(F(k1[q] + k2[q]) + k3[q]) for q in range(4)]
        k[0] = F(k[0] + j["softness"])
        k[3] = F(k[3] + j["softness"])
        a, c, b, d = k[0], k[1], k[2], k[3]  # a = col1.x, b = col2.x, c = col1.y, d = col2.y
        det = F(F(a * d) - F(b * c))
        if det == ZERO:
            return False
        j["m"] = (F(det * d), F(F(-det) * c), F(F(-det) * b), F(det * a))  # quirk Q-3: adjugate TIMES det
        p1, p2 = nn._add(b1.pos, r1), nn._add(b2.pos, r2)
        dp = nn._sub(p2, p1)
        if self.poscorr:
            j["bias"] = nn._scale(nn._scale(nn._scale(dp, inv_dt), j["bias_factor"]), F(-1.0))
        else:
            j["bias"] = _v(0, 0)
        if self.warm:
            p = j["p"]
            b1.vel = nn._sub(b1.vel, nn._scale(p, b1.inv_mass))
            b1.omega = F(b1.omega - F(b1.inv_moi * _cross_vv(r1, p)))
            b2.vel = nn._add(b2.vel, nn._scale(p, b2.inv_mass))
            b2.omega = F(b2.omega + F(b2.inv_moi * _cross_vv(r2, p)))
        else:
            j["p"] = _v(0, 0)
        return True

    def _apply_joint(self, j):  # joint.rs:116-130
        b1, b2 = j["b1"], j["b2"]
        dv = self._dv(b1, b2, j["r1"], j["r2"])
        rhs = nn._sub(nn._sub(j["bias"], dv), nn._scale(j["p"], j["softness"]))
        imp = nn._mul_mv(j["m"], rhs)
        b1.vel = nn._sub(b1.vel, nn._scale(imp, b1.inv_mass))
        b1.omega = F(b1.omega - F(b1.inv_moi * _cross_vv(j["r1"], imp)))
        b2.vel = nn._add(b2.vel, nn._scale(imp, b2.inv_mass))
        b2.omega = F(b2.omega + F(b2.inv_moi * _cross_vv(j["r2"], imp)))
        j["p"] = nn._add(j["p"], imp)

    def step(self, dt):  # world.rs:107-152; returns False when a joint's K had no inverse (the step stops there, quirk Q-15)
        dt = F(dt)
        with np.errstate(all="ignore"):
            inv_dt = F(F(1.0) / dt) if dt > ZERO else ZERO
            self.broad_phase()
            for b in self.bodies:
                if b.inv_mass == ZERO:
                    continue
                b.vel = nn._add(b.vel, nn._scale(nn._add(self.g, nn._scale(b.force, b.inv_mass)), dt))
                b.omega = F(b.omega + F(F(b.inv_moi * b.torque) * dt))
            order = sorted(self.arbiters)
            for key in order:
                self._pre_step_arbiter(self.arbiters[key], inv_dt)
            for j in self.joints:
                if not self._pre_step_joint(j, inv_dt):
                    return False
            for _ in range(self.iterations):
                for key in order:
                    self._apply_arbiter(self.arbiters[key])
                for j in self.joints:
                    self._apply_joint(j)
            for b in self.bodies:
                b.pos = nn._add(b.pos, nn._scale(b.vel, dt))
                b.rot = F(b.rot + F(b.omega * dt))
                b.force, b.torque = _v(0, 0), ZERO
        return True

    def state(self):
        return np.array([(b.pos[0], b.pos[1], b.rot, b.vel[0], b.vel[1], b.omega) for b in self.bodies], dtype=np.float32)
