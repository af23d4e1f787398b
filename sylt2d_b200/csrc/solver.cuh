// solver.cuh — per-constraint device functions of the sequential-impulse solver.
//
//   contact_prestep / contact_apply : Arbiter::pre_step / Arbiter::apply_impulse for ONE contact
//                                     (/root/reference/src/arbiter.rs:182-228, :229-298)
//   joint_prestep / joint_apply     : Joint::pre_step / Joint::apply_impulse
//                                     (/root/reference/src/joint.rs:57-115, :116-130; quirk Q-3 inverse)
//
// f32, reference expression order, no FMA (-fmad=false).  The callers (world.cu: one thread per arbiter of one
// colour; batch.cu: one CTA per world) own the Gauss-Seidel ordering; these functions only do the arithmetic on
// body velocities held in registers.
#pragma once
#include "sylt_math.cuh"

namespace sylt {

struct Vel {
    V2 v;
    float w;
};

struct ContactConst {  // what apply_impulse needs every iteration
    V2 n, r1, r2;
    float mass_n, mass_t, bias;
};

// Arbiter::pre_step for one contact (arbiter.rs:191-223), in two halves so that a caller can compute the constants
// ahead of time (they do not depend on velocities) and keep only the warm start inside the ordered sweep.
//   contact_constants : r1, r2, mass_normal, mass_tangent, bias          (arbiter.rs:192-215)
//   contact_warmstart : apply P = n*pn + t*pt to both bodies when accumulate_impulse (arbiter.rs:216-223, quirk Q-5)
SYLT_HD void contact_constants(V2 cpos, V2 n, float separation, V2 p1, V2 p2, float im1, float ii1, float im2, float ii2,
                               float inv_dt, float k_bias_factor, ContactConst& cc) {
    const float k_allowed_penetration = 0.01f;
    V2 r1 = cpos - p1, r2 = cpos - p2;
    float rn1 = dot(r1, n), rn2 = dot(r2, n);
    float k_normal = im1 + im2;
    k_normal += ii1 * (dot(r1, r1) - rn1 * rn1) + ii2 * (dot(r2, r2) - rn2 * rn2);
    cc.mass_n = 1.0f / k_normal;
    V2 t = cross(n, 1.0f);
    float rt1 = dot(r1, t), rt2 = dot(r2, t);
    float k_tangent = im1 + im2;
    k_tangent += ii1 * (dot(r1, r1) - rt1 * rt1) + ii2 * (dot(r2, r2) - rt2 * rt2);
    cc.mass_t = 1.0f / k_tangent;
    cc.bias = -k_bias_factor * inv_dt * fminf(0.0f, separation + k_allowed_penetration);
    cc.n = n; cc.r1 = r1; cc.r2 = r2;  // apply_impulse recomputes the same r1/r2 (arbiter.rs:236-237): positions do not move during a step
}
SYLT_HD void contact_warmstart(const ContactConst& cc, float im1, float ii1, float im2, float ii2, float pn, float pt, Vel& b1, Vel& b2) {
    V2 t = cross(cc.n, 1.0f);
    V2 p = cc.n * pn + t * pt;
    b1.v = b1.v - p * im1;
    b1.w -= ii1 * cross(cc.r1, p);
    b2.v = b2.v + p * im2;
    b2.w += ii2 * cross(cc.r2, p);
}
SYLT_HD void contact_prestep(V2 cpos, V2 n, float separation, V2 p1, V2 p2, float im1, float ii1, float im2, float ii2,
                             float inv_dt, float k_bias_factor, bool accumulate, float pn, float pt,
                             Vel& b1, Vel& b2, ContactConst& cc) {
    contact_constants(cpos, n, separation, p1, p2, im1, ii1, im2, ii2, inv_dt, k_bias_factor, cc);
    if (accumulate) contact_warmstart(cc, im1, ii1, im2, ii2, pn, pt, b1, b2);
}

// Arbiter::apply_impulse for one contact (arbiter.rs:233-297).  pn/pt are the accumulated impulses (read-modify-write).
SYLT_HD void contact_apply(const ContactConst& c, float friction, float im1, float ii1, float im2, float ii2,
                           bool accumulate, float& pn_acc, float& pt_acc, Vel& b1, Vel& b2) {
    V2 dv = b2.v + cross(b2.w, c.r2) - b1.v - cross(b1.w, c.r1);
    float vn = dot(dv, c.n);
    float d_pn = c.mass_n * (-vn + c.bias);
    if (accumulate) {
        float pn_0 = pn_acc;
        pn_acc = fmaxf(pn_0 + d_pn, 0.0f);
        d_pn = pn_acc - pn_0;
    } else {
        d_pn = fmaxf(0.0f, d_pn);
    }
    V2 pn = c.n * d_pn;
    b1.v = b1.v - pn * im1;
    b1.w -= ii1 * cross(c.r1, pn);
    b2.v = b2.v + pn * im2;
    b2.w += ii2 * cross(c.r2, pn);

    dv = b2.v + cross(b2.w, c.r2) - b1.v - cross(b1.w, c.r1);
    V2 t = cross(c.n, 1.0f);
    float vt = dot(dv, t);
    float d_pt = c.mass_t * -vt;
    if (accumulate) {
        float max_pt = friction * pn_acc;
        float old = pt_acc;
        pt_acc = clampf(old + d_pt, -max_pt, max_pt);
        d_pt = pt_acc - old;
    } else {
        float max_pt = friction * d_pn;
        d_pt = clampf(d_pt, -max_pt, max_pt);
    }
    V2 pt = t * d_pt;
    b1.v = b1.v - pt * im1;
    b1.w -= ii1 * cross(c.r1, pt);
    b2.v = b2.v + pt * im2;
    b2.w += ii2 * cross(c.r2, pt);
}

struct JointConst {
    V2 r1, r2, bias;
    M2 m;
    float softness;
};

// The K matrix of Joint::pre_step (joint.rs:67-94) and whether Mat2x2::invert rejects it (math_utils.rs:190-193).  K depends on
// the poses only, so the joint at which World::step aborts (world.rs:127-129: the FIRST singular joint in Vec order) is known
// before any joint is pre-stepped.
SYLT_HD bool joint_singular(V2 la1, V2 la2, float softness, float c1, float s1, float c2, float s2, float im1, float ii1, float im2, float ii2) {
    V2 r1 = rot_cs(c1, s1) * la1;
    V2 r2 = rot_cs(c2, s2) * la2;
    M2 k1 = mk(mk(im1 + im2, 0.0f), mk(0.0f, im1 + im2));
    M2 k2 = mk(mk(ii1 * r1.y * r1.y, -ii1 * r1.x * r1.y), mk(-ii1 * r1.x * r1.y, ii1 * r1.x * r1.x));
    M2 k3 = mk(mk(ii2 * r2.y * r2.y, -ii2 * r2.x * r2.y), mk(-ii2 * r2.x * r2.y, ii2 * r2.x * r2.x));
    M2 k = k1 + k2 + k3;
    k.c1.x += softness;
    k.c2.y += softness;
    M2 dummy;
    return !invert_q3(k, &dummy, false);  // det == 0 (the opt-in true inverse rejects the same matrices)
}

// Joint::pre_step (joint.rs:57-115).  (c1,s1),(c2,s2): cos/sin of the two body rotations.
// Returns false (and the singular K in *bad) when K has no inverse (quirk Q-15).
SYLT_HD bool joint_prestep(V2 la1, V2 la2, float softness, float bias_factor, V2 p1, float c1, float s1, V2 p2, float c2, float s2,
                           float im1, float ii1, float im2, float ii2, float inv_dt, bool position_correction, bool warm_starting,
                           V2& p_acc, Vel& b1, Vel& b2, JointConst& jc, M2* bad, bool true_inverse = false) {
    V2 r1 = rot_cs(c1, s1) * la1;
    V2 r2 = rot_cs(c2, s2) * la2;
    M2 k1 = mk(mk(im1 + im2, 0.0f), mk(0.0f, im1 + im2));
    M2 k2 = mk(mk(ii1 * r1.y * r1.y, -ii1 * r1.x * r1.y), mk(-ii1 * r1.x * r1.y, ii1 * r1.x * r1.x));
    M2 k3 = mk(mk(ii2 * r2.y * r2.y, -ii2 * r2.x * r2.y), mk(-ii2 * r2.x * r2.y, ii2 * r2.x * r2.x));
    M2 k = k1 + k2 + k3;
    k.c1.x += softness;
    k.c2.y += softness;
    jc.r1 = r1; jc.r2 = r2; jc.softness = softness;
    if (!invert_q3(k, &jc.m, true_inverse)) { *bad = k; return false; }
    V2 a1 = p1 + r1, a2 = p2 + r2;
    V2 dp = a2 - a1;
    if (position_correction) jc.bias = dp * inv_dt * bias_factor * -1.0f;
    else jc.bias = mk(0.0f, 0.0f);
    if (warm_starting) {
        b1.v = b1.v - p_acc * im1;
        b1.w -= ii1 * cross(r1, p_acc);
        b2.v = b2.v + p_acc * im2;
        b2.w += ii2 * cross(r2, p_acc);
    } else {
        p_acc = mk(0.0f, 0.0f);
    }
    return true;
}

// Joint::apply_impulse (joint.rs:116-130)
SYLT_HD void joint_apply(const JointConst& j, float im1, float ii1, float im2, float ii2, V2& p_acc, Vel& b1, Vel& b2) {
    V2 dv = b2.v + cross(b2.w, j.r2) - b1.v - cross(b1.w, j.r1);
    V2 impulse = j.m * (j.bias - dv - p_acc * j.softness);
    b1.v = b1.v - impulse * im1;
    b1.w -= ii1 * cross(j.r1, impulse);
    b2.v = b2.v + impulse * im2;
    b2.w += ii2 * cross(j.r2, impulse);
    p_acc = p_acc + impulse;
}

}  // namespace sylt
