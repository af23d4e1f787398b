// fp_rate.cu — issue rate of non-fused f32 arithmetic on one SM sub-partition (the batch solver is all FMUL/FADD, -fmad=false).
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp_rate fp_rate.cu && ./fp_rate
// Each warp runs 8 independent dependency chains; cycles per warp-instruction are reported for 1, 2 and 4 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8], b = out[1], c = out[2];
    for (int i = 0; i < 8; i++) a[i] = out[3 + i] + threadIdx.x;
    unsigned long long p[8];
    for (int i = 0; i < 8; i++) p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.0f);
    unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) a[i] = a[i] * b;
                if (MODE == 1) a[i] = a[i] + c;
                if (MODE == 2) { if (i & 1) a[i] = a[i] * b; else a[i] = a[i] + c; }
                if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 5) a[i] = fmaxf(a[i], c) ;
                if (MODE == 6) { if (i & 1) a[i] = fmaxf(a[i], b); else a[i] = a[i] + c; }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x + 16] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, float* d, long long* dc) {
    const int iters = 2000;
    for (int threads : {128, 256, 512, 1024}) {
        k<MODE><<<1, threads>>>(d, dc, iters);
        k<MODE><<<1, threads>>>(d, dc, iters);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
        int wps = threads / 128;
        printf("%-24s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", name, wps, (double)c / (iters * 32.0 * wps));
    }
}

int main() {
    float* d; long long* dc;
    cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20); cudaMalloc(&dc, 8);
    run<0>("FMUL", d, dc);
    run<1>("FADD", d, dc);
    run<2>("FMUL/FADD alternating", d, dc);
    run<3>("FMUL2 (f32x2)", d, dc);
    run<4>("FADD2 (f32x2)", d, dc);
    run<5>("FMNMX", d, dc);
    run<6>("FMNMX/FADD alternating", d, dc);
    return 0;
}
