// contact_chain.cu — dependent-chain latency of Arbiter::apply_impulse for one contact (solver.cuh contact_apply), registers only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -I../../sylt2d_b200/csrc -o contact_chain contact_chain.cu
// One warp per SM sub-partition or four; velocities are loop-carried exactly as in a colour hop (contact after contact).
#include <cstdio>
#include <cuda_runtime.h>
#include "solver.cuh"
using namespace sylt;

__global__ void k(const float* in, float* out, long long* cyc, int iters) {
    ContactConst c;
    c.n = mk(in[0], in[1]); c.r1 = mk(in[2] + threadIdx.x * 1e-3f, in[3]); c.r2 = mk(in[4], in[5]);
    c.mass_n = in[6]; c.mass_t = in[7]; c.bias = in[8];
    float fr = in[9], im1 = in[10], ii1 = in[11], im2 = in[12], ii2 = in[13];
    Vel v1, v2;
    v1.v = mk(in[14], in[15]); v1.w = in[16]; v2.v = mk(in[17], in[18]); v2.w = in[19];
    float pn = in[20], pt = in[21];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) contact_apply(c, fr, im1, ii1, im2, ii2, true, pn, pt, v1, v2);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v1.v.x + v1.v.y + v1.w + v2.v.x + v2.v.y + v2.w + pn + pt;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float h[32] = {0.0f, 1.0f, 0.4f, -0.5f, -0.45f, 0.5f, 5.0f, 4.0f, 0.01f, 0.2f, 0.1f, 0.6f, 0.1f, 0.6f, 0.01f, -0.1f, 0.02f, -0.01f, 0.05f, -0.03f, 1.0f, 0.1f};
    float *din, *dout; long long* dc;
    cudaMalloc(&din, sizeof(h)); cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 4096 * 4); cudaMalloc(&dc, 8);
    const int iters = 4000;
    for (int threads : {32, 128, 256, 512}) {
        k<<<1, threads>>>(din, dout, dc, iters);
        k<<<1, threads>>>(din, dout, dc, iters);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d: %.1f cycles per contact_apply\n", threads, (double)c / iters);
    }
    return 0;
}
