// sylt_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.  See sylt_oracle.h.
//
// Literal C++17 restatement of the reference's CPU algorithm for World::step.  Every function cites
// the reference file:line it follows (paths relative to /root/reference).  All arithmetic is IEEE f32
// in the reference's left-to-right expression order; build with -O2 -ffp-contract=off (Rust never
// contracts to FMA).  Reference quirks Q-1..Q-17 of SURVEY.md §0 are reproduced on purpose.
//
// Parity status: "unpinned by reference tests" for step/solver/joints/polygons (the reference has no
// such tests and cannot be built here); pinned only by the box-pair contact counts, math spot values
// and the NoInverse error of the reference's unit tests (tests/test_oracle_kat.py) and cross-checked
// by the independent numpy-f32 restatement oracle/narrow_np.py.
#include "sylt_oracle.h"
#include "../sylt2d_b200/csrc/sylt_sincos.h"  // shared deterministic sin/cos ("device-matched" mode only)

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <thread>
#include <atomic>
#include <utility>
#include <vector>

namespace {

enum { ERR_OK = 0, ERR_NO_INVERSE = 1, ERR_ARBITER = 2, ERR_BODY_ORDER = 3, ERR_BODY_NOT_FOUND = 4,
       ERR_CAPACITY = 5, ERR_CUDA = 6, ERR_INVALID = 7 };

// ---------------------------------------------------------------- math_utils.rs
struct V2 { float x, y; };
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }            // math_utils.rs:108-116
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }            // :129-138
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }             // :118-127
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }                            // :140-149
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }                // :58-60
inline float length(V2 a) { float l = a.x * a.x + a.y * a.y; return sqrtf(l); } // :53-56
inline V2 vabs(V2 a) { return {fabsf(a.x), fabsf(a.y)}; }                     // :62-67
inline float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }              // :81-86
inline V2 cross(V2 a, float s) { return {a.y * s, a.x * -s}; }                // :88-96
inline V2 cross(float s, V2 a) { return {-s * a.y, s * a.x}; }                // :98-106

struct M2 { V2 c1, c2; };
inline V2 operator*(const M2& m, V2 v) {                                       // :230-238
    return {m.c1.x * v.x + m.c2.x * v.y, m.c1.y * v.x + m.c2.y * v.y};
}
inline M2 operator*(const M2& a, const M2& b) { return {a * b.c1, a * b.c2}; } // :240-248
inline M2 operator+(const M2& a, const M2& b) { return {a.c1 + b.c1, a.c2 + b.c2}; } // :219-228
inline M2 transpose(const M2& m) { return {{m.c1.x, m.c2.x}, {m.c1.y, m.c2.y}}; }    // :178-183
inline M2 mabs(const M2& m) { return {vabs(m.c1), vabs(m.c2)}; }               // :201-206

inline M2 rot_from_angle(float a, int mode) {                                  // :164-172
    float c, s;
    if (mode == 0) { c = cosf(a); s = sinf(a); }
    else { c = sylt_cosf(a); s = sylt_sinf(a); }
    return {{c, s}, {-s, c}};
}
// Q-3: multiplies the adjugate by det instead of 1/det (math_utils.rs:185-199)
// fix = true: the opt-in corrected mode (Box2D-lite's Mat22::Invert: scale by 1/det) — NOT the reference's behaviour
inline bool invert(const M2& m, M2* out, bool fix = false) {
    float a = m.c1.x, b = m.c2.x, c = m.c1.y, d = m.c2.y;
    float det = a * d - b * c;
    if (det == 0.0f) return false;
    if (fix) det = 1.0f / det;
    *out = {{det * d, -det * c}, {-det * b, det * a}};
    return true;
}

// ---------------------------------------------------------------- body.rs
struct Poly {
    std::vector<V2> v;
    size_t n() const { return v.size(); }
    // Q-7: shifted by one (body.rs:19-24)
    V2 get_vertex(long i) const { long nn = (long)v.size(); return v[(size_t)((i + nn + 1) % nn)]; }
    V2 get_edge(long i) const { V2 a = get_vertex(i), b = get_vertex(i + 1); return b - a; } // :27-31
    V2 get_normal(long i) const { V2 e = get_edge(i); return {e.y, -e.x}; }                  // :34-41
    V2 centroid() const {                                                                    // :61-81
        float cx = 0.0f, cy = 0.0f, area = 0.0f;
        for (size_t i = 0; i < n(); i++) {
            V2 p1 = get_vertex((long)i), p2 = get_vertex((long)i + 1);
            float cr = p1.x * p2.y - p1.y * p2.x;
            area += cr;
            cx += (p1.x + p2.x) * cr;
            cy += (p1.y + p2.y) * cr;
        }
        area /= 2.0f;
        cx /= 6.0f * area;
        cy /= 6.0f * area;
        return {cx, cy};
    }
    float moi() const {                                                                      // :95-121 (Q-9)
        V2 c = centroid();
        float moi = 0.0f;
        for (size_t i = 0; i < n(); i++) {
            V2 a = get_vertex((long)i), b = get_vertex((long)i + 1);
            V2 p1 = {a.x - c.x, a.y - c.y}, p2 = {b.x - c.x, b.y - c.y};
            float cr = p1.x * p2.y - p1.y * p2.x;
            moi += cr * (p1.x * p1.x + p1.x * p2.x + p2.x * p2.x + p1.y * p1.y + p1.y * p2.y + p2.y * p2.y);
        }
        return fabsf(moi) / 12.0f;
    }
    V2 bounding_box() const {                                                                // :124-146
        float min_x = FLT_MAX, max_x = -FLT_MAX, min_y = FLT_MAX, max_y = -FLT_MAX;
        for (V2 p : v) {
            if (p.x < min_x) min_x = p.x;
            if (p.x > max_x) max_x = p.x;
            if (p.y < min_y) min_y = p.y;
            if (p.y > max_y) max_y = p.y;
        }
        return {max_x - min_x, max_y - min_y};
    }
    Poly rotate(float angle, int mode) const {                                               // :148-158 (Q-12)
        V2 c = centroid();
        M2 r = rot_from_angle(angle, mode);
        Poly o; o.v.reserve(v.size());
        for (V2 p : v) o.v.push_back(r * V2{p.x - c.x, p.y - c.y});
        return o;
    }
    Poly translate(V2 pos) const {                                                           // :160-168
        Poly o; o.v.reserve(v.size());
        for (V2 p : v) o.v.push_back(p + pos);
        return o;
    }
};

struct Body {                                                                                // body.rs:182-199
    uint64_t id = 0;
    V2 position{0, 0}; float rotation = 0;
    V2 velocity{0, 0}; float angular_velocity = 0;
    V2 force{0, 0}; float torque = 0;
    V2 width{0, 0}; float friction = 0, mass = 0, inv_mass = 0, moi = 0, inv_moi = 0;
    std::vector<V2> vertices;
    int shape = 0;  // 0 Box, 1 ConvexPolygon
};

Body body_new_box(V2 width, float mass) {                                                    // body.rs:204-245
    Body b;
    if (mass < FLT_MAX) {
        b.inv_mass = 1.0f / mass;
        b.moi = mass * (width.x * width.x + width.y * width.y) / 12.0f;
        b.inv_moi = 1.0f / b.moi;
    } else { b.inv_mass = 0.0f; b.moi = FLT_MAX; b.inv_moi = 0.0f; }
    float hw = width.x / 2.0f, hh = width.y / 2.0f;
    b.vertices = {{hw, hh}, {-hw, hh}, {-hw, -hh}, {hw, -hh}};
    b.width = width; b.mass = mass; b.shape = 0;
    return b;
}
Body body_new_polygon(const std::vector<V2>& verts, float mass) {                            // body.rs:246-284
    Poly p; p.v = verts;  // Q-8: area() is abs()/2 so orient_counterclockwise never reverses (:44-59)
    Body b;
    if (mass < FLT_MAX) {
        b.inv_mass = 1.0f / mass;
        b.moi = mass * p.moi();
        b.inv_moi = 1.0f / b.moi;
    } else { b.inv_mass = 0.0f; b.moi = FLT_MAX; b.inv_moi = 0.0f; }
    b.width = p.bounding_box();
    b.vertices = verts; b.mass = mass; b.shape = 1;
    return b;
}

// ---------------------------------------------------------------- arbiter.rs types
struct Edges { uint8_t in1 = 0, out1 = 0, in2 = 0, out2 = 0; };                              // arbiter.rs:27-52
struct Feature { Edges e; int32_t value = 0; };                                              // :55-65 (Q-2)
struct Contact {                                                                             // :69-83
    V2 position{0, 0}, normal{0, 0}, r1{0, 0}, r2{0, 0};
    float separation = 0, pn = 0, pt = 0, pnb = 0, mass_normal = 0, mass_tangent = 0, bias = 0;
    Feature feature;
};

// ---------------------------------------------------------------- collide.rs
enum Axis { FaceAX, FaceAY, FaceBX, FaceBY };
struct ClipVertex { Feature fp; V2 v{0, 0}; };                                               // collide.rs:25-38

inline void flip(Feature& fp) { std::swap(fp.e.in1, fp.e.in2); std::swap(fp.e.out1, fp.e.out2); } // :40-43

int clip_segment_to_line(ClipVertex v_out[2], const ClipVertex v_in[2], V2 normal, float offset, uint8_t clip_edge) { // :45-88
    int num_out = 0;
    float d0 = dot(normal, v_in[0].v) - offset;
    float d1 = dot(normal, v_in[1].v) - offset;
    if (d0 <= 0.0f) v_out[num_out++] = v_in[0];
    if (d1 <= 0.0f) v_out[num_out++] = v_in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        v_out[num_out].v = v_in[0].v + (v_in[1].v - v_in[0].v) * interp;
        if (d0 > 0.0f) {
            v_out[num_out].fp = v_in[0].fp;
            v_out[num_out].fp.e.in1 = clip_edge;
            v_out[num_out].fp.e.in2 = 0;
        } else {
            v_out[num_out].fp = v_in[1].fp;
            v_out[num_out].fp.e.out1 = clip_edge;
            v_out[num_out].fp.e.out2 = 0;
        }
        num_out++;
    }
    return num_out;
}

// f32::signum() > 0.0  <=>  sign bit clear and not NaN (collide.rs:100,117)
inline bool signum_pos(float x) { return !std::signbit(x) && !(x != x); }

void compute_incident_edge(ClipVertex c[2], V2 h, V2 pos, const M2& rot, V2 normal) {        // :90-138
    M2 rot_t = transpose(rot);
    V2 n = -(rot_t * normal);
    V2 an = vabs(n);
    ClipVertex c1, c2;
    if (an.x > an.y) {
        if (signum_pos(n.x)) {
            c1.v = {h.x, -h.y}; c1.fp.e.in2 = 3; c1.fp.e.out2 = 4;
            c2.v = {h.x, h.y};  c2.fp.e.in2 = 4; c2.fp.e.out2 = 1;
        } else {
            c1.v = {-h.x, h.y};  c1.fp.e.in2 = 1; c1.fp.e.out2 = 2;
            c2.v = {-h.x, -h.y}; c2.fp.e.in2 = 2; c2.fp.e.out2 = 3;
        }
    } else if (signum_pos(n.y)) {
        c1.v = {h.x, h.y};  c1.fp.e.in2 = 4; c1.fp.e.out2 = 1;
        c2.v = {-h.x, h.y}; c2.fp.e.in2 = 1; c2.fp.e.out2 = 2;
    } else {
        c1.v = {-h.x, -h.y}; c1.fp.e.in2 = 2; c1.fp.e.out2 = 3;
        c2.v = {h.x, -h.y};  c2.fp.e.in2 = 3; c2.fp.e.out2 = 4;
    }
    c1.v = pos + (rot * c1.v);
    c2.v = pos + (rot * c2.v);
    c[0] = c1; c[1] = c2;
}

int collide(std::vector<Contact>& contacts, const Body& A, const Body& B, int mode) {       // :140-297
    V2 h_a = A.width * 0.5f, h_b = B.width * 0.5f;
    V2 pos_a = A.position, pos_b = B.position;
    M2 rot_a = rot_from_angle(A.rotation, mode), rot_b = rot_from_angle(B.rotation, mode);
    M2 rot_a_t = transpose(rot_a), rot_b_t = transpose(rot_b);
    V2 d_p = pos_b - pos_a;
    V2 d_a = rot_a_t * d_p, d_b = rot_b_t * d_p;
    M2 c = rot_a_t * rot_b;
    M2 abs_c = mabs(c), abs_c_t = transpose(mabs(c));
    V2 face_a = vabs(d_a) - h_a - abs_c * h_b;
    if (face_a.x > 0.0f || face_a.y > 0.0f) return 0;
    V2 face_b = vabs(d_b) - h_b - abs_c_t * h_a;
    if (face_b.x > 0.0f || face_b.y > 0.0f) return 0;

    Axis axis = FaceAX;
    float separation = face_a.x;
    V2 normal = d_a.x > 0.0f ? rot_a.c1 : -rot_a.c1;
    const float RELATIVE_TOL = 0.95f, ABSOLUTE_TOL = 0.01f;
    if (face_a.y > RELATIVE_TOL * separation + ABSOLUTE_TOL * h_a.y) {
        axis = FaceAY; separation = face_a.y;
        normal = d_a.y > 0.0f ? rot_a.c2 : -rot_a.c2;
    }
    if (face_b.x > RELATIVE_TOL * separation + ABSOLUTE_TOL * h_b.x) {
        axis = FaceBX; separation = face_b.x;
        normal = d_b.x > 0.0f ? rot_b.c1 : -rot_b.c1;
    }
    if (face_b.y > RELATIVE_TOL * separation + ABSOLUTE_TOL * h_b.y) {
        axis = FaceBY;  // (separation intentionally not updated, :193-196)
        normal = d_b.y > 0.0f ? rot_b.c2 : -rot_b.c2;
    }
    V2 front_normal, side_normal; float neg_side, pos_side, front; uint8_t neg_edge, pos_edge;
    ClipVertex incident[2];
    switch (axis) {
        case FaceAX: {
            front_normal = normal; front = dot(pos_a, front_normal) + h_a.x;
            side_normal = rot_a.c2; float side = dot(pos_a, side_normal);
            neg_side = -side + h_a.y; pos_side = side + h_a.y; neg_edge = 3; pos_edge = 1;
            compute_incident_edge(incident, h_b, pos_b, rot_b, front_normal);
        } break;
        case FaceAY: {
            front_normal = normal; front = dot(pos_a, front_normal) + h_a.y;
            side_normal = rot_a.c1; float side = dot(pos_a, side_normal);
            neg_side = -side + h_a.x; pos_side = side + h_a.x; neg_edge = 2; pos_edge = 4;
            compute_incident_edge(incident, h_b, pos_b, rot_b, front_normal);
        } break;
        case FaceBX: {
            front_normal = -normal; front = dot(pos_b, front_normal) + h_b.x;
            side_normal = rot_b.c2; float side = dot(pos_b, side_normal);
            neg_side = -side + h_b.y; pos_side = side + h_b.y; neg_edge = 3; pos_edge = 1;
            compute_incident_edge(incident, h_a, pos_a, rot_a, front_normal);
        } break;
        default: {
            front_normal = -normal; front = dot(pos_b, front_normal) + h_b.y;
            side_normal = rot_b.c1; float side = dot(pos_b, side_normal);
            neg_side = -side + h_b.x; pos_side = side + h_b.x; neg_edge = 2; pos_edge = 4;
            compute_incident_edge(incident, h_a, pos_a, rot_a, front_normal);
        } break;
    }
    ClipVertex cp1[2], cp2[2];
    int np = clip_segment_to_line(cp1, incident, -side_normal, neg_side, neg_edge);
    if (np < 2) return 0;
    np = clip_segment_to_line(cp2, cp1, side_normal, pos_side, pos_edge);
    if (np < 2) return 0;
    int num = 0;
    for (int i = 0; i < 2; i++) {
        float sep = dot(front_normal, cp2[i].v) - front;
        if (sep <= 0.0f) {
            if (axis == FaceBX || axis == FaceBY) flip(cp2[i].fp);
            Contact ct;
            ct.separation = sep;
            ct.normal = normal;
            ct.position = cp2[i].v - front_normal * sep;
            ct.feature = cp2[i].fp;
            contacts.push_back(ct);
            num++;
        }
    }
    return num;
}

// ---------------------------------------------------------------- collide_polygon.rs
int which_side(const Poly& poly, V2 p, V2 d) {                                               // :18-40
    int pos = 0, neg = 0;
    for (size_t i = 0; i < poly.n(); i++) {
        V2 vtx = poly.get_vertex((long)i);
        float t = dot(d, vtx - p);
        if (t > 0.0f) pos++; else if (t < 0.0f) neg++;
        if (pos > 0 && neg > 0) return 0;
    }
    return pos > 0 ? 1 : -1;
}
bool test_intersection(const Poly& c0, const Poly& c1) {                                     // :51-73
    for (size_t i0 = 0; i0 < c0.n(); i0++) {
        size_t i1 = (i0 + 1) % c0.n();
        V2 p = c0.get_vertex((long)i1), d = c0.get_normal((long)i0);
        if (which_side(c1, p, d) > 0) return false;
    }
    for (size_t i0 = 0; i0 < c1.n(); i0++) {
        size_t i1 = (i0 + 1) % c1.n();
        V2 p = c1.get_vertex((long)i1), d = c1.get_normal((long)i0);
        if (which_side(c0, p, d) > 0) return false;
    }
    return true;
}
std::vector<std::pair<V2, V2>> clip_polygon(const Poly& subject, const Poly& clip) {         // :83-152
    Poly polygon; polygon.v = subject.v;
    std::vector<std::pair<V2, V2>> clipped;
    for (size_t j = 0; j < clip.n(); j++) {
        V2 edge_start = clip.get_vertex((long)j);
        V2 edge_normal = clip.get_normal((long)j);
        std::vector<std::pair<V2, V2>> cur;
        size_t n = polygon.n();
        for (size_t i = 0; i < n; i++) {
            V2 current = polygon.get_vertex((long)i), next = polygon.get_vertex((long)i + 1);
            float dc = dot(edge_normal, current - edge_start) / length(edge_normal);
            float dn = dot(edge_normal, next - edge_start) / length(edge_normal);
            if (dc <= 0.0f) cur.push_back({current, edge_normal});
            if (dc * dn < 0.0f) {
                float interp = dc / (dc - dn);
                V2 inter = current + (next - current) * interp;
                cur.push_back({inter, edge_normal});
            }
        }
        polygon.v.clear();
        for (auto& t : cur) polygon.v.push_back(t.first);
        clipped = cur;
    }
    std::vector<std::pair<V2, V2>> fin;
    for (auto& t : clipped) {
        V2 vertex = t.first;
        V2 closest{0.0f, 0.0f};
        float min_d = FLT_MAX;
        for (size_t j = 0; j < clip.n(); j++) {
            V2 es = clip.get_vertex((long)j), ee = clip.get_vertex((long)j + 1);
            V2 edge = ee - es;
            V2 normal{-edge.y, edge.x};
            normal = normal * (1.0f / length(normal));
            V2 to_point = vertex - es;
            float dist = fabsf(dot(to_point, normal));
            if (dist < min_d) { min_d = dist; closest = normal; }
        }
        fin.push_back({vertex, closest});
    }
    return fin;
}
int collide_polygons(std::vector<Contact>& contacts, const Body& b1, const Body& b2, int mode) { // :164-203
    Poly p1; p1.v = b1.vertices; Poly p2; p2.v = b2.vertices;
    if (p1.n() == 0 || p2.n() == 0) return (int)contacts.size();
    Poly c0 = p1.rotate(b1.rotation, mode).translate(b1.position);
    Poly c1 = p2.rotate(b2.rotation, mode).translate(b2.position);
    if (test_intersection(c0, c1)) {
        auto clipped = clip_polygon(c0, c1);
        contacts.clear();
        for (auto& t : clipped) {
            Contact ct;
            ct.position = t.first;
            ct.normal = t.second;
            ct.separation = dot(t.first, t.second) * 0.001f;  // Q-10
            contacts.push_back(ct);
        }
    }
    return (int)contacts.size();
}

// ---------------------------------------------------------------- arbiter.rs
struct Arbiter {
    int b1 = 0, b2 = 0; float friction = 0; int num_contacts = 0;
    std::vector<Contact> contacts;
    uint64_t stamp = 0;
};
struct Ctx { bool accumulate = true, warm = false, poscorr = true; bool fix_features = false, fix_inverse = false; };                        // world.rs:11-16,38-42 (Q-4)

int narrow(std::vector<Contact>& out, const Body& a, const Body& b, int mode) {              // arbiter.rs:124-127
    if (a.shape == 0 && b.shape == 0) return collide(out, a, b, mode);
    return collide_polygons(out, a, b, mode);
}

void arbiter_update(Arbiter& arb, const std::vector<Contact>& nc, int num_new, const Ctx& ctx) { // :137-181 (Q-2)
    std::vector<Contact> merged;
    for (const Contact& c_new : nc) {
        int k = -1;
        auto packed = [](const Contact& c) { return (uint32_t)c.feature.e.in1 | ((uint32_t)c.feature.e.out1 << 8) | ((uint32_t)c.feature.e.in2 << 16) | ((uint32_t)c.feature.e.out2 << 24); };
        for (size_t j = 0; j < arb.contacts.size(); j++) {
            // reference: FeaturePair.value (never written => always equal, quirk Q-2); corrected mode: the packed edge ids
            bool same = ctx.fix_features ? packed(arb.contacts[j]) == packed(c_new) : arb.contacts[j].feature.value == c_new.feature.value;
            if (same) { k = (int)j; break; }
        }
        if (k > -1) {
            const Contact& c_old = arb.contacts[(size_t)k];
            Contact c = c_new;
            if (ctx.warm) { c.pn = c_old.pn; c.pt = c_old.pt; c.pnb = c_old.pnb; }
            else { c.pn = 0.0f; c.pt = 0.0f; c.pnb = 0.0f; }
            merged.push_back(c);
        } else merged.push_back(c_new);
    }
    arb.contacts = merged;
    arb.num_contacts = num_new;
}

void arbiter_pre_step(Arbiter& arb, std::vector<Body>& bodies, float inv_dt, const Ctx& ctx) {  // :182-228
    const float k_allowed_penetration = 0.01f;
    const float k_bias_factor = ctx.poscorr ? 0.2f : 0.0f;
    Body& body1 = bodies[(size_t)arb.b1]; Body& body2 = bodies[(size_t)arb.b2];
    for (Contact& c : arb.contacts) {
        V2 r1 = c.position - body1.position, r2 = c.position - body2.position;
        float rn1 = dot(r1, c.normal), rn2 = dot(r2, c.normal);
        float k_normal = body1.inv_mass + body2.inv_mass;
        k_normal += body1.inv_moi * (dot(r1, r1) - rn1 * rn1) + body2.inv_moi * (dot(r2, r2) - rn2 * rn2);
        c.mass_normal = 1.0f / k_normal;
        V2 tangent = cross(c.normal, 1.0f);
        float rt1 = dot(r1, tangent), rt2 = dot(r2, tangent);
        float k_tangent = body1.inv_mass + body2.inv_mass;
        k_tangent += body1.inv_moi * (dot(r1, r1) - rt1 * rt1) + body2.inv_moi * (dot(r2, r2) - rt2 * rt2);
        c.mass_tangent = 1.0f / k_tangent;
        c.bias = -k_bias_factor * inv_dt * fminf(0.0f, c.separation + k_allowed_penetration);
        if (ctx.accumulate) {  // Q-5
            V2 p = c.normal * c.pn + tangent * c.pt;
            body1.velocity = body1.velocity - p * body1.inv_mass;
            body1.angular_velocity -= body1.inv_moi * cross(r1, p);
            body2.velocity = body2.velocity + p * body2.inv_mass;
            body2.angular_velocity += body2.inv_moi * cross(r2, p);
        }
    }
}

inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }  // f32::clamp

void arbiter_apply_impulse(Arbiter& arb, std::vector<Body>& bodies, const Ctx& ctx) {         // :229-298
    Body& body1 = bodies[(size_t)arb.b1]; Body& body2 = bodies[(size_t)arb.b2];
    for (Contact& c : arb.contacts) {
        c.r1 = c.position - body1.position;
        c.r2 = c.position - body2.position;
        V2 dv = body2.velocity + cross(body2.angular_velocity, c.r2) - body1.velocity - cross(body1.angular_velocity, c.r1);
        float vn = dot(dv, c.normal);
        float d_pn = c.mass_normal * (-vn + c.bias);
        if (ctx.accumulate) {
            float pn_0 = c.pn;
            c.pn = fmaxf(pn_0 + d_pn, 0.0f);
            d_pn = c.pn - pn_0;
        } else d_pn = fmaxf(0.0f, d_pn);
        V2 pn = c.normal * d_pn;
        body1.velocity = body1.velocity - pn * body1.inv_mass;
        body1.angular_velocity -= body1.inv_moi * cross(c.r1, pn);
        body2.velocity = body2.velocity + pn * body2.inv_mass;
        body2.angular_velocity += body2.inv_moi * cross(c.r2, pn);

        dv = body2.velocity + cross(body2.angular_velocity, c.r2) - body1.velocity - cross(body1.angular_velocity, c.r1);
        V2 tangent = cross(c.normal, 1.0f);
        float vt = dot(dv, tangent);
        float d_pt = c.mass_tangent * -vt;
        if (ctx.accumulate) {
            float max_pt = arb.friction * c.pn;
            float old = c.pt;
            c.pt = clampf(old + d_pt, -max_pt, max_pt);
            d_pt = c.pt - old;
        } else {
            float max_pt = arb.friction * d_pn;
            d_pt = clampf(d_pt, -max_pt, max_pt);
        }
        V2 pt = tangent * d_pt;
        body1.velocity = body1.velocity - pt * body1.inv_mass;
        body1.angular_velocity -= body1.inv_moi * cross(c.r1, pt);
        body2.velocity = body2.velocity + pt * body2.inv_mass;
        body2.angular_velocity += body2.inv_moi * cross(c.r2, pt);
    }
}

// ---------------------------------------------------------------- joint.rs
struct Joint {                                                                               // joint.rs:10-23
    V2 p{0, 0}, bias{0, 0}, r1{0, 0}, r2{0, 0};
    M2 m{{0, 0}, {0, 0}};
    float bias_factor = 0.2f, softness = 0.0f;
    V2 la1{0, 0}, la2{0, 0};
    int b1 = 0, b2 = 0;
};

bool joint_pre_step(Joint& j, std::vector<Body>& bodies, const Ctx& ctx, float inv_dt, int mode, M2* bad) { // :57-115
    Body& body_1 = bodies[(size_t)j.b1]; Body& body_2 = bodies[(size_t)j.b2];
    M2 rot_1 = rot_from_angle(body_1.rotation, mode), rot_2 = rot_from_angle(body_2.rotation, mode);
    j.r1 = rot_1 * j.la1;
    j.r2 = rot_2 * j.la2;
    M2 k1{{0, 0}, {0, 0}};
    k1.c1.x = body_1.inv_mass + body_2.inv_mass; k1.c2.x = 0.0f; k1.c1.y = 0.0f;
    k1.c2.y = body_1.inv_mass + body_2.inv_mass;
    M2 k2{{0, 0}, {0, 0}};
    k2.c1.x = body_1.inv_moi * j.r1.y * j.r1.y;
    k2.c2.x = -body_1.inv_moi * j.r1.x * j.r1.y;
    k2.c1.y = -body_1.inv_moi * j.r1.x * j.r1.y;
    k2.c2.y = body_1.inv_moi * j.r1.x * j.r1.x;
    M2 k3{{0, 0}, {0, 0}};
    k3.c1.x = body_2.inv_moi * j.r2.y * j.r2.y;
    k3.c2.x = -body_2.inv_moi * j.r2.x * j.r2.y;
    k3.c1.y = -body_2.inv_moi * j.r2.x * j.r2.y;
    k3.c2.y = body_2.inv_moi * j.r2.x * j.r2.x;
    M2 k = k1 + k2 + k3;
    k.c1.x += j.softness;
    k.c2.y += j.softness;
    if (!invert(k, &j.m, ctx.fix_inverse)) { *bad = k; return false; }  // Q-15: mid-step abort
    V2 p1 = body_1.position + j.r1, p2 = body_2.position + j.r2;
    V2 dp = p2 - p1;
    if (ctx.poscorr) j.bias = dp * inv_dt * j.bias_factor * -1.0f;
    else j.bias = {0.0f, 0.0f};
    if (ctx.warm) {
        body_1.velocity = body_1.velocity - j.p * body_1.inv_mass;
        body_1.angular_velocity -= body_1.inv_moi * cross(j.r1, j.p);
        body_2.velocity = body_2.velocity + j.p * body_2.inv_mass;
        body_2.angular_velocity += body_2.inv_moi * cross(j.r2, j.p);
    } else j.p = {0.0f, 0.0f};
    return true;
}
void joint_apply_impulse(Joint& j, std::vector<Body>& bodies) {                              // :116-130
    Body& body_1 = bodies[(size_t)j.b1]; Body& body_2 = bodies[(size_t)j.b2];
    V2 dv = body_2.velocity + cross(body_2.angular_velocity, j.r2) - body_1.velocity - cross(body_1.angular_velocity, j.r1);
    V2 impulse = j.m * (j.bias - dv - j.p * j.softness);
    body_1.velocity = body_1.velocity - impulse * body_1.inv_mass;
    body_1.angular_velocity -= body_1.inv_moi * cross(j.r1, impulse);
    body_2.velocity = body_2.velocity + impulse * body_2.inv_mass;
    body_2.angular_velocity += body_2.inv_moi * cross(j.r2, impulse);
    j.p = j.p + impulse;
}

}  // namespace

// ---------------------------------------------------------------- world.rs
struct OrcWorld {
    V2 gravity{0, 0};
    uint32_t iterations = 10;
    Ctx ctx;
    std::vector<Body> bodies;
    std::vector<Joint> joints;
    // Q-1: the reference uses a HashMap with random iteration order; the canonical order here is
    // ascending (body1_id, body2_id) (arbiter.rs:85-105), Box2D-lite's std::map order.
    std::map<std::pair<uint64_t, uint64_t>, Arbiter> arbiters;
    int sincos_mode = 0, broadphase = 0;
    uint64_t stamp = 0;
    M2 bad{{0, 0}, {0, 0}};
    int bad_joint = -1;

    int visit_pair(size_t i, size_t j) {                                                     // world.rs:78-101
        const Body& bi = bodies[i]; const Body& bj = bodies[j];
        if (bi.inv_mass == 0.0f && bj.inv_mass == 0.0f) return ERR_OK;
        if (bi.id > bj.id) return ERR_BODY_ORDER;  // Q-6: reference panics in RefCell::swap (arbiter.rs:120-122)
        std::vector<Contact> nc; nc.reserve(2);
        int num = narrow(nc, bi, bj, sincos_mode);
        auto key = std::make_pair(bi.id, bj.id);
        if (num > 0) {
            auto it = arbiters.find(key);
            if (it != arbiters.end()) {
                arbiter_update(it->second, nc, num, ctx);
                it->second.stamp = stamp;
            } else {
                Arbiter a; a.b1 = (int)i; a.b2 = (int)j;
                a.friction = sqrtf(bi.friction * bj.friction);                               // arbiter.rs:128 (Q-17)
                a.num_contacts = num; a.contacts = nc; a.stamp = stamp;
                arbiters.emplace(key, std::move(a));
            }
        } else if (broadphase == 0) arbiters.erase(key);
        return ERR_OK;
    }

    void body_aabb(const Body& b, float* lo, float* hi) const {
        // conservative world AABB (grid mode only; not part of the reference)
        M2 r = rot_from_angle(b.rotation, sincos_mode);
        float minx = FLT_MAX, maxx = -FLT_MAX, miny = FLT_MAX, maxy = -FLT_MAX;
        auto acc = [&](V2 p) { minx = std::min(minx, p.x); maxx = std::max(maxx, p.x); miny = std::min(miny, p.y); maxy = std::max(maxy, p.y); };
        if (b.shape == 0) {
            V2 h = b.width * 0.5f;
            const float sx[4] = {1, -1, -1, 1}, sy[4] = {1, 1, -1, -1};
            for (int k = 0; k < 4; k++) acc(b.position + r * V2{h.x * sx[k], h.y * sy[k]});
        }
        if (b.shape == 1 || true) {
            Poly p; p.v = b.vertices;
            if (p.n()) { V2 c = p.centroid(); for (V2 q : p.v) acc(b.position + r * V2{q.x - c.x, q.y - c.y}); }
        }
        float m = 1e-4f * (fabsf(b.position.x) + fabsf(b.position.y) + (maxx - minx) + (maxy - miny)) + 1e-5f;
        lo[0] = minx - m; lo[1] = miny - m; hi[0] = maxx + m; hi[1] = maxy + m;
    }

    int broad_phase_grid() {
        size_t n = bodies.size();
        std::vector<float> lo(2 * n), hi(2 * n);
        std::vector<float> ext; ext.reserve(n);
        for (size_t i = 0; i < n; i++) {
            body_aabb(bodies[i], &lo[2 * i], &hi[2 * i]);
            if (bodies[i].inv_mass != 0.0f) ext.push_back(std::max(hi[2 * i] - lo[2 * i], hi[2 * i + 1] - lo[2 * i + 1]));
        }
        std::vector<std::pair<uint32_t, uint32_t>> cand;
        if (ext.empty()) return ERR_OK;
        std::vector<float> se = ext; std::nth_element(se.begin(), se.begin() + se.size() / 2, se.end());
        float med = se[se.size() / 2];
        float cell = 0.0f;
        for (float e : ext) if (e <= 4.0f * med) cell = std::max(cell, e);
        if (!(cell > 0.0f)) cell = 1.0f;
        std::vector<uint32_t> small, large;
        for (size_t i = 0; i < n; i++) {
            float e = std::max(hi[2 * i] - lo[2 * i], hi[2 * i + 1] - lo[2 * i + 1]);
            if (e <= cell) small.push_back((uint32_t)i); else large.push_back((uint32_t)i);
        }
        auto overlap = [&](size_t a, size_t b) {
            return !(lo[2 * a] > hi[2 * b] || lo[2 * b] > hi[2 * a] || lo[2 * a + 1] > hi[2 * b + 1] || lo[2 * b + 1] > hi[2 * a + 1]);
        };
        struct Ent { int64_t cx, cy; uint32_t i; };
        std::vector<Ent> ents; ents.reserve(small.size());
        for (uint32_t i : small) {
            float cx = 0.5f * (lo[2 * i] + hi[2 * i]), cy = 0.5f * (lo[2 * i + 1] + hi[2 * i + 1]);
            ents.push_back({(int64_t)floorf(cx / cell), (int64_t)floorf(cy / cell), i});
        }
        std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.cx != b.cx ? a.cx < b.cx : (a.cy != b.cy ? a.cy < b.cy : a.i < b.i); });
        auto lower = [&](int64_t cx, int64_t cy) {
            return std::lower_bound(ents.begin(), ents.end(), Ent{cx, cy, 0}, [](const Ent& a, const Ent& b) { return a.cx != b.cx ? a.cx < b.cx : a.cy < b.cy; });
        };
        for (const Ent& e : ents)
            for (int64_t dx = -1; dx <= 1; dx++)
                for (int64_t dy = -1; dy <= 1; dy++)
                    for (auto it = lower(e.cx + dx, e.cy + dy); it != ents.end() && it->cx == e.cx + dx && it->cy == e.cy + dy; ++it)
                        if (it->i > e.i && overlap(e.i, it->i)) cand.push_back({e.i, it->i});
        for (uint32_t L : large)
            for (size_t k = 0; k < n; k++) {
                if (k == L) continue;
                bool k_large = std::max(hi[2 * k] - lo[2 * k], hi[2 * k + 1] - lo[2 * k + 1]) > cell;
                if (k_large && k < L) continue;  // large-large handled once
                if (overlap(L, k)) cand.push_back({(uint32_t)std::min<size_t>(L, k), (uint32_t)std::max<size_t>(L, k)});
            }
        std::sort(cand.begin(), cand.end());
        cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
        for (auto& c : cand) { int e = visit_pair(c.first, c.second); if (e) return e; }
        return ERR_OK;
    }

    int broad_phase() {                                                                      // world.rs:73-105
        stamp++;
        if (broadphase == 0) {
            for (size_t i = 0; i < bodies.size(); i++)
                for (size_t j = i + 1; j < bodies.size(); j++) { int e = visit_pair(i, j); if (e) return e; }
            return ERR_OK;
        }
        int e = broad_phase_grid();
        if (e) return e;
        for (auto it = arbiters.begin(); it != arbiters.end();) { if (it->second.stamp != stamp) it = arbiters.erase(it); else ++it; }
        return ERR_OK;
    }

    int step(float dt, const uint64_t* keys, int n_arb, const int32_t* jorder, int n_joint) { // world.rs:107-152
        float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
        int e = broad_phase();
        if (e) return e;
        for (Body& b : bodies) {                                                             // :113-120
            if (b.inv_mass == 0.0f) continue;
            b.velocity = b.velocity + (gravity + b.force * b.inv_mass) * dt;
            b.angular_velocity += b.inv_moi * b.torque * dt;
        }
        std::vector<Arbiter*> order;
        if (keys) {
            if ((size_t)n_arb != arbiters.size()) return ERR_INVALID;
            order.reserve((size_t)n_arb);
            for (int k = 0; k < n_arb; k++) {
                auto it = arbiters.find({keys[2 * k], keys[2 * k + 1]});
                if (it == arbiters.end()) return ERR_INVALID;
                order.push_back(&it->second);
            }
            std::vector<Arbiter*> chk = order; std::sort(chk.begin(), chk.end());
            if (std::adjacent_find(chk.begin(), chk.end()) != chk.end()) return ERR_INVALID;
        } else {
            order.reserve(arbiters.size());
            for (auto& kv : arbiters) order.push_back(&kv.second);
        }
        std::vector<int> jo;
        if (jorder) {
            if ((size_t)n_joint != joints.size()) return ERR_INVALID;
            jo.assign(jorder, jorder + n_joint);
            for (int j : jo) if (j < 0 || (size_t)j >= joints.size()) return ERR_INVALID;
        } else { jo.resize(joints.size()); for (size_t j = 0; j < jo.size(); j++) jo[j] = (int)j; }

        for (Arbiter* a : order) arbiter_pre_step(*a, bodies, inv_dt, ctx);                  // :123-125
        // :127-129  `for joint in self.joints.iter_mut() { joint.pre_step(..)?; }` returns at the FIRST joint in Vec order whose K
        // has no inverse; joints behind it are never touched.  K depends on the poses only (joint.rs:67-94), so under an explicit
        // replay order the same joints are pre-stepped (in the replayed order) and the same joint reports the error.
        int jfail = 0x7fffffff;
        if (jorder)
            for (size_t j = 0; j < joints.size(); j++) {
                Joint probe = joints[j];
                std::vector<Body> two = {bodies[(size_t)probe.b1], bodies[(size_t)probe.b2]};
                probe.b1 = 0; probe.b2 = 1;
                M2 k;
                if (!joint_pre_step(probe, two, ctx, inv_dt, sincos_mode, &k)) { jfail = (int)j; break; }
            }
        for (int j : jo) {
            if (j > jfail) continue;
            if (!joint_pre_step(joints[(size_t)j], bodies, ctx, inv_dt, sincos_mode, &bad)) {
                bad_joint = j;
                if (!jorder) return ERR_NO_INVERSE;  // canonical order == Vec order: the reference's loop, literally
            }
        }
        if (jfail != 0x7fffffff) return ERR_NO_INVERSE;
        for (uint32_t it = 0; it < iterations; it++) {                                       // :132-140
            for (Arbiter* a : order) arbiter_apply_impulse(*a, bodies, ctx);
            for (int j : jo) joint_apply_impulse(joints[(size_t)j], bodies);
        }
        for (Body& b : bodies) {                                                             // :143-150 (Q-16)
            b.position = b.position + b.velocity * dt;
            b.rotation += b.angular_velocity * dt;
            b.force = {0.0f, 0.0f};
            b.torque = 0.0f;
        }
        return ERR_OK;
    }

    int find_body(uint64_t id) const {                                                       // joint.rs:27-36 (first match)
        for (size_t i = 0; i < bodies.size(); i++) if (bodies[i].id == id) return (int)i;
        return -1;
    }
};

// ================================================================ C ABI
extern "C" {

OrcWorld* orc_world_create(float gx, float gy, uint32_t iterations) {                        // world.rs:37-51
    OrcWorld* w = new OrcWorld();
    w->gravity = {gx, gy};
    w->iterations = iterations;
    return w;
}
void orc_world_destroy(OrcWorld* w) { delete w; }
void orc_world_clear(OrcWorld* w) { w->bodies.clear(); w->joints.clear(); w->arbiters.clear(); } // world.rs:67-71
void orc_world_set_context(OrcWorld* w, int a, int ws, int pc) { w->ctx.accumulate = a != 0; w->ctx.warm = ws != 0; w->ctx.poscorr = pc != 0; }
void orc_world_set_mode(OrcWorld* w, int sincos_mode, int broadphase) { w->sincos_mode = sincos_mode; w->broadphase = broadphase; }
void orc_world_set_iterations(OrcWorld* w, uint32_t it) { w->iterations = it; }
void orc_world_set_fixes(OrcWorld* w, int fix_features, int fix_inverse) { w->ctx.fix_features = fix_features != 0; w->ctx.fix_inverse = fix_inverse != 0; }

static int push_body(OrcWorld* w, Body b, uint64_t id, float friction, float x, float y, float rot, float vx, float vy, float om) {
    b.id = id; b.friction = friction; b.position = {x, y}; b.rotation = rot; b.velocity = {vx, vy}; b.angular_velocity = om;
    w->bodies.push_back(std::move(b));
    return (int)w->bodies.size() - 1;
}
int orc_world_add_box(OrcWorld* w, uint64_t id, float wx, float wy, float mass, float friction,
                      float x, float y, float rot, float vx, float vy, float omega) {
    return push_body(w, body_new_box({wx, wy}, mass), id, friction, x, y, rot, vx, vy, omega);
}
int orc_world_add_polygon(OrcWorld* w, uint64_t id, const float* v, int n, float mass, float friction,
                          float x, float y, float rot, float vx, float vy, float omega) {
    if (n < 1) return -ERR_INVALID;
    std::vector<V2> vs; for (int i = 0; i < n; i++) vs.push_back({v[2 * i], v[2 * i + 1]});
    return push_body(w, body_new_polygon(vs, mass), id, friction, x, y, rot, vx, vy, omega);
}
int orc_world_add_joint(OrcWorld* w, uint64_t id1, uint64_t id2, float ax, float ay) {       // joint.rs:26-55
    int b1 = w->find_body(id1), b2 = w->find_body(id2);
    if (b1 < 0 || b2 < 0) return -ERR_BODY_NOT_FOUND;
    Joint j; j.b1 = b1; j.b2 = b2;
    M2 rt1 = transpose(rot_from_angle(w->bodies[(size_t)b1].rotation, w->sincos_mode));
    M2 rt2 = transpose(rot_from_angle(w->bodies[(size_t)b2].rotation, w->sincos_mode));
    V2 anchor{ax, ay};
    j.la1 = rt1 * (anchor - w->bodies[(size_t)b1].position);
    j.la2 = rt2 * (anchor - w->bodies[(size_t)b2].position);
    w->joints.push_back(j);
    return (int)w->joints.size() - 1;
}
int orc_world_set_joint_params(OrcWorld* w, int j, float softness, float bias_factor) {
    if (j < 0 || (size_t)j >= w->joints.size()) return ERR_INVALID;
    w->joints[(size_t)j].softness = softness; w->joints[(size_t)j].bias_factor = bias_factor;
    return ERR_OK;
}
int orc_world_add_force(OrcWorld* w, int b, float fx, float fy) {                            // body.rs:286-288
    if (b < 0 || (size_t)b >= w->bodies.size()) return ERR_INVALID;
    w->bodies[(size_t)b].force = w->bodies[(size_t)b].force + V2{fx, fy};
    return ERR_OK;
}
int orc_world_num_bodies(OrcWorld* w) { return (int)w->bodies.size(); }
int orc_world_num_joints(OrcWorld* w) { return (int)w->joints.size(); }
int orc_world_num_arbiters(OrcWorld* w) { return (int)w->arbiters.size(); }
int orc_world_num_contacts(OrcWorld* w) { int c = 0; for (auto& kv : w->arbiters) c += (int)kv.second.contacts.size(); return c; }

int orc_world_get_bodies(OrcWorld* w, OrcBodyState* out, int cap) {
    int n = std::min<int>(cap, (int)w->bodies.size());
    for (int i = 0; i < n; i++) {
        const Body& b = w->bodies[(size_t)i];
        out[i] = {b.position.x, b.position.y, b.rotation, b.velocity.x, b.velocity.y, b.angular_velocity, b.force.x, b.force.y, b.torque};
    }
    return n;
}
int orc_world_set_bodies(OrcWorld* w, int first, int count, const OrcBodyState* in) {
    if (first < 0 || count < 0 || (size_t)(first + count) > w->bodies.size()) return ERR_INVALID;
    for (int i = 0; i < count; i++) {
        Body& b = w->bodies[(size_t)(first + i)];
        b.position = {in[i].x, in[i].y}; b.rotation = in[i].rotation; b.velocity = {in[i].vx, in[i].vy};
        b.angular_velocity = in[i].angular_velocity; b.force = {in[i].fx, in[i].fy}; b.torque = in[i].torque;
    }
    return ERR_OK;
}
int orc_world_get_body_props(OrcWorld* w, float* o, int cap) {
    int n = std::min<int>(cap, (int)w->bodies.size());
    for (int i = 0; i < n; i++) {
        const Body& b = w->bodies[(size_t)i];
        float* p = o + 8 * i;
        p[0] = b.mass; p[1] = b.inv_mass; p[2] = b.moi; p[3] = b.inv_moi; p[4] = b.width.x; p[5] = b.width.y; p[6] = b.friction; p[7] = (float)b.shape;
    }
    return n;
}
// the pub Body fields the step reads every time (inv_mass / inv_moi: world.rs:115-118, arbiter.rs, joint.rs; width: collide.rs:141-142;
// friction: arbiter.rs:128 at arbiter creation): a caller of the reference may overwrite them between steps
int orc_world_set_body_props(OrcWorld* w, int first, int count, const float* in) {
    if (first < 0 || count < 0 || (size_t)first + (size_t)count > w->bodies.size()) return ERR_INVALID;
    for (int i = 0; i < count; i++) {
        Body& b = w->bodies[(size_t)(first + i)];
        const float* p = in + 8 * i;
        b.mass = p[0]; b.inv_mass = p[1]; b.moi = p[2]; b.inv_moi = p[3]; b.width = {p[4], p[5]}; b.friction = p[6];
    }
    return ERR_OK;
}
// pub Joint.local_anchor_1 / local_anchor_2 (joint.rs:19-20)
int orc_world_set_joint_local_anchors(OrcWorld* w, int j, float la1x, float la1y, float la2x, float la2y) {
    if (j < 0 || (size_t)j >= w->joints.size()) return ERR_INVALID;
    w->joints[(size_t)j].la1 = {la1x, la1y}; w->joints[(size_t)j].la2 = {la2x, la2y};
    return ERR_OK;
}
static void export_contact(const Contact& c, OrcContact* o) {
    o->px = c.position.x; o->py = c.position.y; o->nx = c.normal.x; o->ny = c.normal.y;
    o->r1x = c.r1.x; o->r1y = c.r1.y; o->r2x = c.r2.x; o->r2y = c.r2.y;
    o->separation = c.separation; o->pn = c.pn; o->pt = c.pt; o->pnb = c.pnb;
    o->mass_normal = c.mass_normal; o->mass_tangent = c.mass_tangent; o->bias = c.bias;
    o->in_edge_1 = c.feature.e.in1; o->out_edge_1 = c.feature.e.out1; o->in_edge_2 = c.feature.e.in2; o->out_edge_2 = c.feature.e.out2;
    o->feature_value = c.feature.value;
}
int orc_world_get_arbiters(OrcWorld* w, OrcArbiterInfo* arb, int arb_cap, OrcContact* contacts, int contact_cap) {
    int a = 0, c = 0;
    for (auto& kv : w->arbiters) {
        if (a >= arb_cap) return -ERR_CAPACITY;
        const Arbiter& ar = kv.second;
        if (c + (int)ar.contacts.size() > contact_cap) return -ERR_CAPACITY;
        arb[a] = {kv.first.first, kv.first.second, ar.b1, ar.b2, ar.friction, ar.num_contacts, c, 0};
        for (const Contact& ct : ar.contacts) export_contact(ct, &contacts[c++]);
        a++;
    }
    return a;
}
int orc_world_get_joints(OrcWorld* w, OrcJointState* out, int cap) {
    int n = std::min<int>(cap, (int)w->joints.size());
    for (int i = 0; i < n; i++) {
        const Joint& j = w->joints[(size_t)i];
        out[i] = {j.p.x, j.p.y, j.bias.x, j.bias.y, j.r1.x, j.r1.y, j.r2.x, j.r2.y, j.m.c1.x, j.m.c2.x, j.m.c1.y, j.m.c2.y,
                  j.bias_factor, j.softness, j.la1.x, j.la1.y, j.la2.x, j.la2.y, j.b1, j.b2};
    }
    return n;
}
int orc_world_broad_phase(OrcWorld* w) { return w->broad_phase(); }
int orc_world_step(OrcWorld* w, float dt) { return w->step(dt, nullptr, 0, nullptr, 0); }
int orc_world_step_ordered(OrcWorld* w, float dt, const uint64_t* keys, int n_arb, const int32_t* jo, int n_joint) {
    return w->step(dt, keys, n_arb, jo, n_joint);
}
int orc_world_last_error_matrix(OrcWorld* w, float* o) {
    o[0] = w->bad.c1.x; o[1] = w->bad.c2.x; o[2] = w->bad.c1.y; o[3] = w->bad.c2.y;
    return w->bad_joint;
}

int orc_collide_pair(int mode,
                     int shape_a, float wax, float way, const float* va, int na, float xa, float ya, float ra,
                     int shape_b, float wbx, float wby, const float* vb, int nb, float xb, float yb, float rb,
                     OrcContact* out, int cap) {
    auto mk = [](int shape, float wx, float wy, const float* v, int n) {
        if (shape == 0) return body_new_box({wx, wy}, 1.0f);
        std::vector<V2> vs; for (int i = 0; i < n; i++) vs.push_back({v[2 * i], v[2 * i + 1]});
        return body_new_polygon(vs, 1.0f);
    };
    Body A = mk(shape_a, wax, way, va, na), B = mk(shape_b, wbx, wby, vb, nb);
    A.position = {xa, ya}; A.rotation = ra; B.position = {xb, yb}; B.rotation = rb;
    std::vector<Contact> cs;
    int n = narrow(cs, A, B, mode);
    if (n > cap) return -ERR_CAPACITY;
    for (int i = 0; i < n; i++) export_contact(cs[(size_t)i], &out[i]);
    return n;
}

int orc_mat_invert(const float* m, float* o) {
    M2 in{{m[0], m[2]}, {m[1], m[3]}}, out;  // m = (col1.x, col2.x, col1.y, col2.y)
    if (!invert(in, &out)) return ERR_NO_INVERSE;
    o[0] = out.c1.x; o[1] = out.c2.x; o[2] = out.c1.y; o[3] = out.c2.y;
    return ERR_OK;
}
void orc_sincos(int mode, float a, float* s, float* c) {
    if (mode == 0) { *c = cosf(a); *s = sinf(a); } else { *c = sylt_cosf(a); *s = sylt_sinf(a); }
}

int orc_worlds_step_mt(OrcWorld** worlds, int n_worlds, float dt, int steps, int threads) {
    if (threads < 1) threads = 1;
    std::atomic<int> next{0}, err{0};
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_worlds) break;
            for (int s = 0; s < steps; s++) { int e = worlds[i]->step(dt, nullptr, 0, nullptr, 0); if (e) { err = e; break; } }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return err.load();
}

}  // extern "C"
