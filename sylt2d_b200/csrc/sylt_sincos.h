// sylt_sincos.h — one deterministic f32 sin/cos shared by host and device code.
//
// Why it exists: the reference evaluates `f32::cos` / `f32::sin` inside Mat2x2::new_from_angle
// (/root/reference/src/math_utils.rs:164-172) for every pair in collide() (src/collide.rs:146-147),
// every polygon rotate (src/body.rs:148-158), Joint::new (src/joint.rs:37-38) and every joint
// pre_step (src/joint.rs:64-65).  Rust lowers f32::sin/cos to the platform libm (glibc sinf/cosf on
// Linux).  CUDA's sinf/cosf are not bit-identical to glibc's, so a pair whose SAT face value sits
// within an ulp of zero could flip between host and device and break the bit-exact pair-set gate.
//
// What it is: a restatement of the published single-precision algorithm glibc >= 2.28 uses
// (ARM "optimized routines" sinf/cosf: double-precision reduction by pi/2, degree-7/8 minimax
// polynomials, 4/pi bit table for |x| >= 120), written with EXPLICIT fma() at exactly the places
// where the FMA-enabled x86-64 glibc build contracts, and nowhere else.  It uses only IEEE double
// * + fma, integer ops and double<->int conversions, which are bit-identical on x86-64 and on
// sm_100a (compile with nvcc -fmad=false / gcc -ffp-contract=off so nothing else is contracted).
// tests/test_sincos.py checks it against the host libm over the whole f32 range (sampled in the
// fast suite, exhaustive under SYLT_SLOW=1): on glibc 2.39 / x86-64 with FMA the mismatch count is 0,
// i.e. "device-matched" mode == "reference-faithful" mode on such hosts.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SYLT_HD __host__ __device__ __forceinline__
#else
#define SYLT_HD inline
#endif

namespace sylt_sc {

// Polynomial coefficients of the two quadrant tables (second one yields -cos).
struct Tab {
    double c0, c1, c2, c3, c4, s1, s2, s3;
};

SYLT_HD uint32_t asuint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
SYLT_HD uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

SYLT_HD Tab table(int neg) {
    Tab t;
    t.c0 = 0x1p0;
    t.c1 = -0x1.ffffffd0c621cp-2;
    t.c2 = 0x1.55553e1068f19p-5;
    t.c3 = -0x1.6c087e89a359dp-10;
    t.c4 = 0x1.99343027bf8c3p-16;
    if (neg) {
        t.c0 = -t.c0; t.c1 = -t.c1; t.c2 = -t.c2; t.c3 = -t.c3; t.c4 = -t.c4;
    }
    t.s1 = -0x1.555545995a603p-3;
    t.s2 = 0x1.1107605230bc4p-7;
    t.s3 = -0x1.994eb3774cf24p-13;
    return t;
}

SYLT_HD double sign_of(int n) {  // {1,-1,-1,1}[n & 3]
    int k = n & 3;
    return (k == 1 || k == 2) ? -1.0 : 1.0;
}

// sin polynomial when n is even, cos polynomial when n is odd.
SYLT_HD float poly(double x, double x2, const Tab& p, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = fma(x2, p.s3, p.s2);
        double x7 = x3 * x2;
        double s = fma(x3, p.s1, x);
        return (float)fma(x7, s1, s);
    } else {
        double x4 = x2 * x2;
        double c2 = fma(x2, p.c4, p.c3);
        double c1 = fma(x2, p.c1, p.c0);
        double x6 = x4 * x2;
        double c = fma(x4, p.c2, c1);
        return (float)fma(x6, c2, c);
    }
}

SYLT_HD double reduce_fast(double x, int* np) {
    const double hpi_inv = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
    const double hpi = 0x1.921FB54442D18p0;
    double r = x * hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, hpi, x);
}

SYLT_HD uint32_t inv_pio4(int i) {  // bits of 4/pi, 8-bit stride windows of 32 bits
    switch (i) {
        case 0: return 0xa2u;        case 1: return 0xa2f9u;      case 2: return 0xa2f983u;
        case 3: return 0xa2f9836eu;  case 4: return 0xf9836e4eu;  case 5: return 0x836e4e44u;
        case 6: return 0x6e4e4415u;  case 7: return 0x4e441529u;  case 8: return 0x441529fcu;
        case 9: return 0x1529fc27u;  case 10: return 0x29fc2757u; case 11: return 0xfc2757d1u;
        case 12: return 0x2757d1f5u; case 13: return 0x57d1f534u; case 14: return 0xd1f534ddu;
        case 15: return 0xf534ddc0u; case 16: return 0x34ddc0dbu; case 17: return 0xddc0db62u;
        case 18: return 0xc0db6295u; case 19: return 0xdb629599u; case 20: return 0x6295993cu;
        case 21: return 0x95993c43u; case 22: return 0x993c4390u; default: return 0x3c439041u;
    }
}

SYLT_HD double reduce_large(uint32_t xi, int* np) {
    int base = (xi >> 26) & 15;
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * inv_pio4(base));
    res1 = (uint64_t)xi * inv_pio4(base + 4);
    res2 = (uint64_t)xi * inv_pio4(base + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * 0x1.921FB54442D18p-62;
}

}  // namespace sylt_sc

SYLT_HD float sylt_sinf(float y) {
    using namespace sylt_sc;
    double x = (double)y;
    int n;
    if (abstop12(y) < 0x3f4u) {  // abstop12(pi/4)
        double s = x * x;
        if (abstop12(y) < 0x398u) return y;  // |y| < 2^-12
        return poly(x, s, table(0), 0);
    } else if (abstop12(y) < 0x42fu) {  // |y| < 120
        x = reduce_fast(x, &n);
        double s = sign_of(n);
        return poly(x * s, x * x, table(n & 2), n);
    } else if (abstop12(y) < 0x7f8u) {
        uint32_t xi = asuint(y);
        int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        double s = sign_of(n + sign);
        return poly(x * s, x * x, table((n + sign) & 2), n);
    }
    return y - y;  // inf/nan -> nan
}

SYLT_HD float sylt_cosf(float y) {
    using namespace sylt_sc;
    double x = (double)y;
    int n;
    if (abstop12(y) < 0x3f4u) {
        double x2 = x * x;
        if (abstop12(y) < 0x398u) return 1.0f;
        return poly(x, x2, table(0), 1);
    } else if (abstop12(y) < 0x42fu) {
        x = reduce_fast(x, &n);
        double s = sign_of(n);
        return poly(x * s, x * x, table(n & 2), n ^ 1);
    } else if (abstop12(y) < 0x7f8u) {
        uint32_t xi = asuint(y);
        int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        double s = sign_of(n + sign);
        return poly(x * s, x * x, table((n + sign) & 2), n ^ 1);
    }
    return y - y;
}

// (cos, sin) of one angle, as Mat2x2::new_from_angle evaluates them (cos first, then sin).
SYLT_HD void sylt_sincosf(float a, float* s, float* c) {
    *c = sylt_cosf(a);
    *s = sylt_sinf(a);
}
