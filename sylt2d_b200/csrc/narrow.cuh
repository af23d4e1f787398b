// narrow.cuh — per-pair narrow phase, one thread per candidate pair.
//
//   box_box   : OBB SAT over 4 axes with the reference's relative/absolute axis preference, incident-edge
//               selection and two segment clips; <= 2 contacts with feature edges.
//               Restates /root/reference/src/collide.rs:40-297.
//   poly_poly : SAT over all edges of both polygons, Sutherland-Hodgman clip of c0 by c1, nearest-edge unit
//               normals; variable contact count (quirk Q-11).  Restates src/collide_polygon.rs:18-203 together
//               with the shifted vertex accessor of src/body.rs:19-41 (quirk Q-7).
//
// Everything is f32 in the reference's expression order (compile with -fmad=false).  Feature edges are packed
// one byte each into a u32: in_edge_1 | out_edge_1<<8 | in_edge_2<<16 | out_edge_2<<24.
#pragma once
#include "sylt_math.cuh"

namespace sylt {

struct ContactOut {
    float px, py, nx, ny, sep;
    uint32_t feat;
};

constexpr int MAX_POLY_VERTS = 32;               // per polygon (SYLT_MAX_POLY_VERTS); the kernels come in tiers of 8 / 16 / 32 vertices,
constexpr int MAX_MANIFOLD = 2 * MAX_POLY_VERTS; // chosen by the largest polygon of the world; a manifold has at most n0 + n1 contacts

struct ClipV {
    V2 v;
    uint32_t fp;
};

// clip_segment_to_line (collide.rs:45-88)
SYLT_HD int clip_segment(ClipV* vout, const ClipV* vin, V2 normal, float offset, uint32_t clip_edge) {
    int n = 0;
    vout[0].fp = 0; vout[0].v = mk(0.0f, 0.0f);
    vout[1].fp = 0; vout[1].v = mk(0.0f, 0.0f);
    float d0 = dot(normal, vin[0].v) - offset;
    float d1 = dot(normal, vin[1].v) - offset;
    if (d0 <= 0.0f) vout[n++] = vin[0];
    if (d1 <= 0.0f) vout[n++] = vin[1];
    if (d0 * d1 < 0.0f) {
        float t = d0 / (d0 - d1);
        ClipV o;
        o.v = vin[0].v + (vin[1].v - vin[0].v) * t;
        if (d0 > 0.0f) o.fp = (vin[0].fp & 0xff00ff00u) | clip_edge;              // in_edge_1 = clip, in_edge_2 = none
        else o.fp = (vin[1].fp & 0x00ff00ffu) | (clip_edge << 8);                  // out_edge_1 = clip, out_edge_2 = none
        vout[n++] = o;
    }
    return n;
}

// compute_incident_edge (collide.rs:90-138): the edge of the other box most anti-parallel to the face normal
SYLT_HD void incident_edge(ClipV* c, V2 h, V2 pos, M2 rot, V2 normal) {
    V2 n = -(transpose(rot) * normal);
    V2 an = vabs(n);
    V2 a, b;
    uint32_t fa, fb;  // (in_edge_2, out_edge_2) packed at bits 16..31
    if (an.x > an.y) {
        if (signum_pos(n.x)) { a = mk(h.x, -h.y); fa = (3u << 16) | (4u << 24); b = mk(h.x, h.y);   fb = (4u << 16) | (1u << 24); }
        else                 { a = mk(-h.x, h.y); fa = (1u << 16) | (2u << 24); b = mk(-h.x, -h.y); fb = (2u << 16) | (3u << 24); }
    } else if (signum_pos(n.y)) { a = mk(h.x, h.y);   fa = (4u << 16) | (1u << 24); b = mk(-h.x, h.y); fb = (1u << 16) | (2u << 24); }
    else                        { a = mk(-h.x, -h.y); fa = (2u << 16) | (3u << 24); b = mk(h.x, -h.y); fb = (3u << 16) | (4u << 24); }
    c[0].v = pos + (rot * a); c[0].fp = fa;
    c[1].v = pos + (rot * b); c[1].fp = fb;
}

// collide (collide.rs:140-297).  (ca,sa),(cb,sb) = (cos,sin) of the two rotations; wa, wb = Body.width.
SYLT_HD int box_box(V2 pos_a, float ca, float sa, V2 wa, V2 pos_b, float cb, float sb, V2 wb, ContactOut* out) {
    V2 h_a = wa * 0.5f, h_b = wb * 0.5f;
    M2 rot_a = rot_cs(ca, sa), rot_b = rot_cs(cb, sb);
    M2 rot_a_t = transpose(rot_a), rot_b_t = transpose(rot_b);
    V2 d_p = pos_b - pos_a;
    V2 d_a = rot_a_t * d_p;
    V2 d_b = rot_b_t * d_p;
    M2 c = rot_a_t * rot_b;
    M2 abs_c = mabs(c);
    M2 abs_c_t = transpose(abs_c);
    V2 face_a = vabs(d_a) - h_a - abs_c * h_b;
    if (face_a.x > 0.0f || face_a.y > 0.0f) return 0;
    V2 face_b = vabs(d_b) - h_b - abs_c_t * h_a;
    if (face_b.x > 0.0f || face_b.y > 0.0f) return 0;

    // axis of least penetration with the 0.95 / 0.01 preference (collide.rs:171-196)
    int axis = 0;  // 0 FaceAX, 1 FaceAY, 2 FaceBX, 3 FaceBY
    float separation = face_a.x;
    V2 normal = d_a.x > 0.0f ? rot_a.c1 : -rot_a.c1;
    const float REL = 0.95f, ABS = 0.01f;
    if (face_a.y > REL * separation + ABS * h_a.y) { axis = 1; separation = face_a.y; normal = d_a.y > 0.0f ? rot_a.c2 : -rot_a.c2; }
    if (face_b.x > REL * separation + ABS * h_b.x) { axis = 2; separation = face_b.x; normal = d_b.x > 0.0f ? rot_b.c1 : -rot_b.c1; }
    if (face_b.y > REL * separation + ABS * h_b.y) { axis = 3; normal = d_b.y > 0.0f ? rot_b.c2 : -rot_b.c2; }

    // reference-face frame (collide.rs:206-251)
    const bool on_b = axis >= 2;
    const bool y_face = (axis & 1) != 0;
    V2 front_normal = on_b ? -normal : normal;
    V2 rp = on_b ? pos_b : pos_a;
    V2 rh = on_b ? h_b : h_a;
    M2 rr = on_b ? rot_b : rot_a;
    float front = dot(rp, front_normal) + (y_face ? rh.y : rh.x);
    V2 side_normal = y_face ? rr.c1 : rr.c2;
    float side = dot(rp, side_normal);
    float hs = y_face ? rh.x : rh.y;
    float neg_side = -side + hs, pos_side = side + hs;
    uint32_t neg_edge = y_face ? 2u : 3u, pos_edge = y_face ? 4u : 1u;
    ClipV inc[2], cp1[2], cp2[2];
    if (on_b) incident_edge(inc, h_a, pos_a, rot_a, front_normal);
    else incident_edge(inc, h_b, pos_b, rot_b, front_normal);

    if (clip_segment(cp1, inc, -side_normal, neg_side, neg_edge) < 2) return 0;
    if (clip_segment(cp2, cp1, side_normal, pos_side, pos_edge) < 2) return 0;
    int n = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float sep = dot(front_normal, cp2[i].v) - front;
        if (sep <= 0.0f) {
            uint32_t fp = cp2[i].fp;
            if (on_b) fp = (fp >> 16) | (fp << 16);  // flip (collide.rs:40-43)
            V2 p = cp2[i].v - front_normal * sep;
            out[n].px = p.x; out[n].py = p.y; out[n].nx = normal.x; out[n].ny = normal.y; out[n].sep = sep; out[n].feat = fp;
            n++;
        }
    }
    return n;
}

// ------------------------------------------------------------------ polygons
// ConvexPolygon::get_vertex with its +1 shift (body.rs:19-24, quirk Q-7)
// (callers pass 0 <= i <= n: conditional wraps instead of an integer modulo, the same index)
SYLT_HD V2 pv(const V2* w, int n, int i) { int k = i + 1; if (k >= n) k -= n; if (k >= n) k -= n; return w[k]; }  // (second wrap: a clipped polygon of one vertex)
// get_normal (body.rs:34-41): (e.y, -e.x) of edge i
SYLT_HD V2 pnormal(const V2* w, int n, int i) {
    V2 e = pv(w, n, i + 1) - pv(w, n, i);
    return mk(e.y, -e.x);
}
// which_side (collide_polygon.rs:18-40)
SYLT_HD int which_side(const V2* w, int n, V2 p, V2 d) {
    int pos = 0, neg = 0;
    for (int i = 0; i < n; i++) {
        float t = dot(d, pv(w, n, i) - p);
        if (t > 0.0f) pos++;
        else if (t < 0.0f) neg++;
        if (pos > 0 && neg > 0) return 0;
    }
    return pos > 0 ? 1 : -1;
}
// test_intersection (collide_polygon.rs:51-73)
SYLT_HD bool polys_intersect(const V2* c0, int n0, const V2* c1, int n1) {
    for (int i0 = 0; i0 < n0; i0++) {
        int i1 = (i0 + 1) % n0;
        if (which_side(c1, n1, pv(c0, n0, i1), pnormal(c0, n0, i0)) > 0) return false;
    }
    for (int i0 = 0; i0 < n1; i0++) {
        int i1 = (i0 + 1) % n1;
        if (which_side(c0, n0, pv(c1, n1, i1), pnormal(c1, n1, i0)) > 0) return false;
    }
    return true;
}

// collide_polygons (collide_polygon.rs:83-203) on world-space vertex lists c0 (body1) and c1 (body2).
// Returns the contact count; *overflow is set when the clipped polygon would exceed 2 * MAXV vertices (never for convex input).
template <int MAXV>  // vertices per polygon the caller's tier allows: the clipped polygon has at most 2 * MAXV
SYLT_HD int poly_poly(const V2* c0, int n0, const V2* c1, int n1, ContactOut* out, bool* overflow) {
    constexpr int MAX_MANIFOLD = 2 * MAXV;
    if (n0 <= 0 || n1 <= 0) return 0;
    if (!polys_intersect(c0, n0, c1, n1)) return 0;
    V2 bufa[MAX_MANIFOLD], bufb[MAX_MANIFOLD];
    V2* poly = bufa;
    V2* cur = bufb;
    int np = n0;
    for (int i = 0; i < n0; i++) poly[i] = c0[i];
    int nc = 0;
    for (int j = 0; j < n1; j++) {  // clip_polygon (collide_polygon.rs:90-124)
        V2 es = pv(c1, n1, j);
        V2 en = pnormal(c1, n1, j);
        float el = len(en);
        nc = 0;
        for (int i = 0; i < np; i++) {
            V2 a = pv(poly, np, i), b = pv(poly, np, i + 1);
            float dc = dot(en, a - es) / el;
            float dn = dot(en, b - es) / el;
            if (dc <= 0.0f) {
                if (nc < MAX_MANIFOLD) cur[nc] = a; else *overflow = true;
                nc++;
            }
            if (dc * dn < 0.0f) {
                float t = dc / (dc - dn);
                if (nc < MAX_MANIFOLD) cur[nc] = a + (b - a) * t; else *overflow = true;
                nc++;
            }
        }
        if (nc > MAX_MANIFOLD) nc = MAX_MANIFOLD;
        V2* tmp = poly; poly = cur; cur = tmp;
        np = nc;
    }
    // nearest-edge unit normal of c1 for every surviving vertex (collide_polygon.rs:127-149), contacts (:175-192)
    V2 unit[MAXV];  // the unit normal of an edge does not depend on the vertex: once per edge (the reference recomputes it per vertex, same bits)
    for (int j = 0; j < n1; j++) {
        V2 e = pv(c1, n1, j + 1) - pv(c1, n1, j);
        V2 nrm = mk(-e.y, e.x);
        unit[j] = nrm * (1.0f / len(nrm));
    }
    for (int k = 0; k < np; k++) {
        V2 vtx = poly[k];
        V2 best = mk(0.0f, 0.0f);
        float md = 3.402823466e+38f;
        for (int j = 0; j < n1; j++) {
            V2 nrm = unit[j];
            float d = fabsf(dot(vtx - pv(c1, n1, j), nrm));
            if (d < md) { md = d; best = nrm; }
        }
        out[k].px = vtx.x; out[k].py = vtx.y; out[k].nx = best.x; out[k].ny = best.y;
        out[k].sep = dot(vtx, best) * 0.001f;  // quirk Q-10
        out[k].feat = 0;
    }
    return np;
}

}  // namespace sylt
