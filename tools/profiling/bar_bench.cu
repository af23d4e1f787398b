#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, int mode) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 0) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory"); }
        else { __threadfence(); atomicAdd(counter, 1); }
        unsigned v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
        if (mode != 2) __threadfence();
    }
    __syncthreads();
}
__global__ void k_bar(unsigned* counter, int n, int mode, float4* vel, int nb, int work) {
    unsigned target = 0;
    cg::grid_group grid = cg::this_grid();
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < n; i++) {
        if (work && threadIdx.x < work) {
            unsigned h = (tid * 2654435761u + i * 40503u);
            int b1 = h % nb, b2 = (h >> 7) % nb;
            float4 a = __ldcg(&vel[b1]), b = __ldcg(&vel[b2]);
            #pragma unroll 1
            for (int k = 0; k < 60; k++) { a.x = a.x * 1.0001f + b.y; b.y = b.y * 0.9999f + a.x; }
            __stcg(&vel[b1], a); __stcg(&vel[b2], b);
        }
        if (mode == 3) grid.sync(); else grid_barrier(counter, target, mode);
    }
}
int main() {
    unsigned* c; float4* vel; int nb = 100000;
    cudaMalloc(&c, 4); cudaMalloc(&vel, nb * 16); cudaMemset(vel, 0, nb * 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {512, 128}) for (int work : {0, 128}) for (int mode = 0; mode < 4; mode++) {
        if (work > threads) continue;
        int n = 2000;
        void* args[] = {&c, &n, &mode, &vel, &nb, &work};
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(c, 0, 4);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void*)k_bar, dim3(sms), dim3(threads), args, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("threads %d work %d mode %d (%s): %.3f us per barrier  err=%s\n", threads, work, mode,
               mode == 0 ? "red.release+fence" : mode == 1 ? "fence+atomicAdd+fence" : mode == 2 ? "fence+atomicAdd, no acquire fence" : "cg grid.sync",
               ms * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
