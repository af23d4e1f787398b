// sincos_check — test infrastructure (not product): compares the shared sylt_sinf/sylt_cosf
// (sylt2d_b200/csrc/sylt_sincos.h) with the host libm sinf/cosf, i.e. with what Rust's
// f32::sin / f32::cos evaluate to on this platform (/root/reference/src/math_utils.rs:165-166).
// usage: sincos_check <stride> [threads]   (stride 1 = all 2^32 bit patterns)
#include "../sylt2d_b200/csrc/sylt_sincos.h"
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
int main(int argc, char** argv) {
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 257;
    int nt = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::atomic<uint64_t> total{0}, ms{0}, mc{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            uint64_t n = 0, a = 0, b = 0;
            for (uint64_t bits = (uint64_t)t * stride; bits < (1ULL << 32); bits += stride * nt) {
                uint32_t u = (uint32_t)bits;
                float x;
                memcpy(&x, &u, 4);
                float s = sylt_sinf(x), c = sylt_cosf(x);
                float ls = sinf(x), lc = cosf(x);
                n++;
                if (memcmp(&s, &ls, 4) != 0 && !(s != s && ls != ls)) a++;
                if (memcmp(&c, &lc, 4) != 0 && !(c != c && lc != lc)) b++;
            }
            total += n; ms += a; mc += b;
        });
    for (auto& t : th) t.join();
    printf("{\"checked\": %llu, \"sin_mismatch\": %llu, \"cos_mismatch\": %llu}\n",
           (unsigned long long)total.load(), (unsigned long long)ms.load(), (unsigned long long)mc.load());
    return 0;
}
