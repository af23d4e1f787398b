"""Second, independent restatement of the reference's two narrow phases in numpy-float32 scalars.

TEST INFRASTRUCTURE, NOT PRODUCT.  Purpose (SURVEY.md §8c "Mitigation"): the reference's own tests only pin
contact COUNTS, and the Rust crate cannot be built here, so the C++ oracle (oracle/sylt_oracle.cpp) is
cross-validated bit-for-bit against this separately written version.  Pure-Python loops: small cases only.

Follows /root/reference/src/collide.rs:40-297 (boxes) and src/collide_polygon.rs:18-203 + src/body.rs:14-172
(polygons).  sin/cos are injected (`sincos(angle) -> (s, c)`) so both restatements share one libm.
"""
import numpy as np

F = np.float32
ZERO = F(0.0)


def _mul_mv(m, v):  # Mat2x2 * Vec2 (math_utils.rs:230-238); m = (c1x, c1y, c2x, c2y)
    return (F(F(m[0] * v[0]) + F(m[2] * v[1])), F(F(m[1] * v[0]) + F(m[3] * v[1])))


def _rot(angle, sincos):  # math_utils.rs:164-172
    s, c = sincos(F(angle))
    s, c = F(s), F(c)
    return (c, s, F(-s), c)


def _T(m):  # transpose (math_utils.rs:178-183)
    return (m[0], m[2], m[1], m[3])


def _dot(a, b):
    return F(F(a[0] * b[0]) + F(a[1] * b[1]))


def _sub(a, b):
    return (F(a[0] - b[0]), F(a[1] - b[1]))


def _add(a, b):
    return (F(a[0] + b[0]), F(a[1] + b[1]))


def _scale(a, s):
    return (F(a[0] * s), F(a[1] * s))


def _neg(a):
    return (F(-a[0]), F(-a[1]))


def _sig_pos(x):  # f32::signum() > 0.0
    return (not np.signbit(x)) and not np.isnan(x)


def _clip(v_in, normal, offset, clip_edge):
    """clip_segment_to_line (collide.rs:45-88). vertices: dict(v=(x,y), e=[in1,out1,in2,out2])."""
    out = []
    d0 = F(_dot(normal, v_in[0]["v"]) - offset)
    d1 = F(_dot(normal, v_in[1]["v"]) - offset)
    if d0 <= ZERO:
        out.append({"v": v_in[0]["v"], "e": list(v_in[0]["e"])})
    if d1 <= ZERO:
        out.append({"v": v_in[1]["v"], "e": list(v_in[1]["e"])})
    if F(d0 * d1) < ZERO:
        t = F(d0 / F(d0 - d1))
        p = _add(v_in[0]["v"], _scale(_sub(v_in[1]["v"], v_in[0]["v"]), t))
        if d0 > ZERO:
            e = list(v_in[0]["e"])
            e[0] = clip_edge
            e[2] = 0
        else:
            e = list(v_in[1]["e"])
            e[1] = clip_edge
            e[3] = 0
        out.append({"v": p, "e": e})
    return out


def _incident(h, pos, rot, normal):
    """compute_incident_edge (collide.rs:90-138)."""
    n = _neg(_mul_mv(_T(rot), normal))
    ax, ay = F(abs(n[0])), F(abs(n[1]))
    if ax > ay:
        if _sig_pos(n[0]):
            a, b = ((h[0], F(-h[1])), [0, 0, 3, 4]), ((h[0], h[1]), [0, 0, 4, 1])
        else:
            a, b = ((F(-h[0]), h[1]), [0, 0, 1, 2]), ((F(-h[0]), F(-h[1])), [0, 0, 2, 3])
    elif _sig_pos(n[1]):
        a, b = ((h[0], h[1]), [0, 0, 4, 1]), ((F(-h[0]), h[1]), [0, 0, 1, 2])
    else:
        a, b = ((F(-h[0]), F(-h[1])), [0, 0, 2, 3]), ((h[0], F(-h[1])), [0, 0, 3, 4])
    return [{"v": _add(pos, _mul_mv(rot, a[0])), "e": a[1]}, {"v": _add(pos, _mul_mv(rot, b[0])), "e": b[1]}]


def collide_boxes(wa, pa, ra, wb, pb, rb, sincos):
    """collide (collide.rs:140-297). Returns list of dict(sep, normal, pos, edges, axis)."""
    wa, pa, wb, pb = [(F(t[0]), F(t[1])) for t in (wa, pa, wb, pb)]
    hA, hB = _scale(wa, F(0.5)), _scale(wb, F(0.5))
    RA, RB = _rot(ra, sincos), _rot(rb, sincos)
    RAt, RBt = _T(RA), _T(RB)
    dp = _sub(pb, pa)
    dA, dB = _mul_mv(RAt, dp), _mul_mv(RBt, dp)
    c1, c2 = _mul_mv(RAt, (RB[0], RB[1])), _mul_mv(RAt, (RB[2], RB[3]))
    absC = (F(abs(c1[0])), F(abs(c1[1])), F(abs(c2[0])), F(abs(c2[1])))
    fA = _sub(_sub((F(abs(dA[0])), F(abs(dA[1]))), hA), _mul_mv(absC, hB))
    if fA[0] > ZERO or fA[1] > ZERO:
        return []
    fB = _sub(_sub((F(abs(dB[0])), F(abs(dB[1]))), hB), _mul_mv(_T(absC), hA))
    if fB[0] > ZERO or fB[1] > ZERO:
        return []
    A1, A2, B1, B2 = (RA[0], RA[1]), (RA[2], RA[3]), (RB[0], RB[1]), (RB[2], RB[3])
    axis, sep = "AX", fA[0]
    normal = A1 if dA[0] > ZERO else _neg(A1)
    REL, ABS = F(0.95), F(0.01)
    if fA[1] > F(F(REL * sep) + F(ABS * hA[1])):
        axis, sep = "AY", fA[1]
        normal = A2 if dA[1] > ZERO else _neg(A2)
    if fB[0] > F(F(REL * sep) + F(ABS * hB[0])):
        axis, sep = "BX", fB[0]
        normal = B1 if dB[0] > ZERO else _neg(B1)
    if fB[1] > F(F(REL * sep) + F(ABS * hB[1])):
        axis = "BY"
        normal = B2 if dB[1] > ZERO else _neg(B2)
    if axis == "AX":
        fn = normal
        front = F(_dot(pa, fn) + hA[0])
        sn = A2
        side = _dot(pa, sn)
        neg_side, pos_side, ne, pe = F(F(-side) + hA[1]), F(side + hA[1]), 3, 1
        inc = _incident(hB, pb, RB, fn)
    elif axis == "AY":
        fn = normal
        front = F(_dot(pa, fn) + hA[1])
        sn = A1
        side = _dot(pa, sn)
        neg_side, pos_side, ne, pe = F(F(-side) + hA[0]), F(side + hA[0]), 2, 4
        inc = _incident(hB, pb, RB, fn)
    elif axis == "BX":
        fn = _neg(normal)
        front = F(_dot(pb, fn) + hB[0])
        sn = B2
        side = _dot(pb, sn)
        neg_side, pos_side, ne, pe = F(F(-side) + hB[1]), F(side + hB[1]), 3, 1
        inc = _incident(hA, pa, RA, fn)
    else:
        fn = _neg(normal)
        front = F(_dot(pb, fn) + hB[1])
        sn = B1
        side = _dot(pb, sn)
        neg_side, pos_side, ne, pe = F(F(-side) + hB[0]), F(side + hB[0]), 2, 4
        inc = _incident(hA, pa, RA, fn)
    c1v = _clip(inc, _neg(sn), neg_side, ne)
    if len(c1v) < 2:
        return []
    c2v = _clip(c1v, sn, pos_side, pe)
    if len(c2v) < 2:
        return []
    out = []
    for cv in c2v[:2]:
        s = F(_dot(fn, cv["v"]) - front)
        if s <= ZERO:
            e = list(cv["e"])
            if axis in ("BX", "BY"):
                e = [e[2], e[3], e[0], e[1]]
            out.append({"sep": s, "normal": normal, "pos": _sub(cv["v"], _scale(fn, s)), "edges": tuple(e), "axis": axis})
    return out


# ------------------------------------------------------------------ polygons
def _gv(v, i):  # ConvexPolygon::get_vertex with the +1 shift (body.rs:19-24, quirk Q-7)
    n = len(v)
    return v[(i + n + 1) % n]


def _centroid(v):  # body.rs:61-81
    cx = cy = area = ZERO
    for i in range(len(v)):
        p1, p2 = _gv(v, i), _gv(v, i + 1)
        cr = F(F(p1[0] * p2[1]) - F(p1[1] * p2[0]))
        area = F(area + cr)
        cx = F(cx + F(F(p1[0] + p2[0]) * cr))
        cy = F(cy + F(F(p1[1] + p2[1]) * cr))
    area = F(area / F(2.0))
    with np.errstate(all="ignore"):
        cx = F(cx / F(F(6.0) * area))
        cy = F(cy / F(F(6.0) * area))
    return (cx, cy)


def world_polygon(verts, pos, rot, sincos):  # rotate + translate (body.rs:148-168, quirk Q-12)
    v = [(F(p[0]), F(p[1])) for p in verts]
    c = _centroid(v)
    R = _rot(rot, sincos)
    pos = (F(pos[0]), F(pos[1]))
    return [_add(_mul_mv(R, (F(p[0] - c[0]), F(p[1] - c[1]))), pos) for p in v]


def _normal(v, i):  # get_normal (body.rs:34-41)
    e = _sub(_gv(v, i + 1), _gv(v, i))
    return (e[1], F(-e[0]))


def _which_side(poly, p, d):  # collide_polygon.rs:18-40
    pos = neg = 0
    for i in range(len(poly)):
        t = _dot(d, _sub(_gv(poly, i), p))
        if t > ZERO:
            pos += 1
        elif t < ZERO:
            neg += 1
        if pos > 0 and neg > 0:
            return 0
    return 1 if pos > 0 else -1


def _intersects(c0, c1):  # collide_polygon.rs:51-73
    for a, b in ((c0, c1), (c1, c0)):
        for i0 in range(len(a)):
            i1 = (i0 + 1) % len(a)
            if _which_side(b, _gv(a, i1), _normal(a, i0)) > 0:
                return False
    return True


def _len(a):
    return F(np.sqrt(F(F(a[0] * a[0]) + F(a[1] * a[1]))))


def collide_polys(va, pa, ra, vb, pb, rb, sincos):
    """collide_polygons (collide_polygon.rs:195-203). Returns list of dict(sep, normal, pos)."""
    c0 = world_polygon(va, pa, ra, sincos)
    c1 = world_polygon(vb, pb, rb, sincos)
    if not _intersects(c0, c1):
        return []
    poly = list(c0)
    clipped = []
    with np.errstate(all="ignore"):
        for j in range(len(c1)):  # clip_polygon (collide_polygon.rs:83-124)
            es, en = _gv(c1, j), _normal(c1, j)
            cur = []
            n = len(poly)
            for i in range(n):
                a, b = _gv(poly, i), _gv(poly, i + 1)
                dc = F(_dot(en, _sub(a, es)) / _len(en))
                dn = F(_dot(en, _sub(b, es)) / _len(en))
                if dc <= ZERO:
                    cur.append(a)
                if F(dc * dn) < ZERO:
                    t = F(dc / F(dc - dn))
                    cur.append(_add(a, _scale(_sub(b, a), t)))
            poly = cur
            clipped = cur
        out = []
        for vtx in clipped:  # nearest-edge normal (collide_polygon.rs:127-149)
            best, md = (ZERO, ZERO), F(np.finfo(np.float32).max)
            for j in range(len(c1)):
                es, ee = _gv(c1, j), _gv(c1, j + 1)
                e = _sub(ee, es)
                nrm = (F(-e[1]), e[0])
                nrm = _scale(nrm, F(F(1.0) / _len(nrm)))
                d = F(abs(_dot(_sub(vtx, es), nrm)))
                if d < md:
                    md, best = d, nrm
            out.append({"sep": F(_dot(vtx, best) * F(0.001)), "normal": best, "pos": vtx, "edges": (0, 0, 0, 0)})
    return out


BOX_VERTS = lambda w: [(F(w[0]) / F(2.0), F(w[1]) / F(2.0)), (-F(w[0]) / F(2.0), F(w[1]) / F(2.0)),
                       (-F(w[0]) / F(2.0), -F(w[1]) / F(2.0)), (F(w[0]) / F(2.0), -F(w[1]) / F(2.0))]  # body.rs:217-224
