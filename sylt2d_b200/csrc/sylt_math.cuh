// sylt_math.cuh — f32 vector/matrix helpers with the reference's exact expression order.
// Restates /root/reference/src/math_utils.rs:53-106,164-248 as __host__ __device__ inline code.
// Build with nvcc -fmad=false (and -Xcompiler -ffp-contract=off for the host side) so that every product is
// rounded before it is added, as in the Rust reference, which never contracts to FMA.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "sylt_sincos.h"

namespace sylt {

struct V2 {
    float x, y;
};
SYLT_HD V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
SYLT_HD V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }   // math_utils.rs:108-116
SYLT_HD V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }   // :129-138
SYLT_HD V2 operator*(V2 a, float s) { return mk(a.x * s, a.y * s); }    // :118-127
SYLT_HD V2 operator-(V2 a) { return mk(-a.x, -a.y); }                   // :140-149
SYLT_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }         // :58-60
SYLT_HD float len(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }        // :53-56
SYLT_HD V2 vabs(V2 a) { return mk(fabsf(a.x), fabsf(a.y)); }            // :62-67
SYLT_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }       // :81-86
SYLT_HD V2 cross(V2 a, float s) { return mk(a.y * s, a.x * -s); }       // :88-96
SYLT_HD V2 cross(float s, V2 a) { return mk(-s * a.y, s * a.x); }       // :98-106

struct M2 {
    V2 c1, c2;  // columns
};
SYLT_HD M2 mk(V2 c1, V2 c2) { M2 m; m.c1 = c1; m.c2 = c2; return m; }
SYLT_HD V2 operator*(M2 m, V2 v) {                                       // :230-238
    return mk(m.c1.x * v.x + m.c2.x * v.y, m.c1.y * v.x + m.c2.y * v.y);
}
SYLT_HD M2 operator*(M2 a, M2 b) { return mk(a * b.c1, a * b.c2); }      // :240-248
SYLT_HD M2 operator+(M2 a, M2 b) { return mk(a.c1 + b.c1, a.c2 + b.c2); }  // :219-228
SYLT_HD M2 transpose(M2 m) { return mk(mk(m.c1.x, m.c2.x), mk(m.c1.y, m.c2.y)); }  // :178-183
SYLT_HD M2 mabs(M2 m) { return mk(vabs(m.c1), vabs(m.c2)); }             // :201-206
// Mat2x2::new_from_angle from an already evaluated (cos, sin)           // :164-172
SYLT_HD M2 rot_cs(float c, float s) { return mk(mk(c, s), mk(-s, c)); }
SYLT_HD M2 rot_angle(float a) { return rot_cs(sylt_cosf(a), sylt_sinf(a)); }

// Mat2x2::invert as the reference has it (quirk Q-3: adjugate * det, not / det)   // :185-199
// true_inverse = true is the opt-in corrected mode (scale by 1/det, Box2D-lite's Mat22::Invert), not the reference.
SYLT_HD bool invert_q3(M2 m, M2* out, bool true_inverse = false) {
    float a = m.c1.x, b = m.c2.x, c = m.c1.y, d = m.c2.y;
    float det = a * d - b * c;
    if (det == 0.0f) return false;
    if (true_inverse) det = 1.0f / det;
    *out = mk(mk(det * d, -det * c), mk(-det * b, det * a));
    return true;
}

// f32::signum() > 0.0: sign bit clear and not NaN (collide.rs:100,117)
SYLT_HD bool signum_pos(float x) {
#if defined(__CUDA_ARCH__)
    return !(__float_as_uint(x) >> 31) && !(x != x);
#else
    return !signbit(x) && !(x != x);
#endif
}
// f32::clamp
SYLT_HD float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

}  // namespace sylt
