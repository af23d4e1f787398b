// L2 ping-pong latency between CTAs with the solver's versioned-record protocol (relaxed gpu-scope 64-bit ld/st).
// chain of N hops over P participants (CTAs on different SMs): participant p waits for version k where k % P == p.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ldr(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void str(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__global__ void k_chain(unsigned long long* cell, int hops, int lanes, int work) {
    // every lane of warp 0 runs an independent chain on its own cell (stride 64 B); the warp moves in lock step
    int P = gridDim.x, p = blockIdx.x, lane = threadIdx.x;
    if (lane >= lanes) return;
    unsigned long long* c = cell + lane * 8;
    float x = 1.0f;
    for (int k = p; k < hops; k += P) {
        while ((unsigned)(ldr(c) >> 32) != (unsigned)k) {}
        for (int q = 0; q < work; q++) x = x * 1.0001f + 0.5f;   // dependent flops standing in for contact_apply
        str(c, ((unsigned long long)(k + 1) << 32) | __float_as_uint(x));
    }
}
int main() {
    unsigned long long* cell; cudaMalloc(&cell, 32 * 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int P : {2, 8, 148}) for (int lanes : {1, 32}) for (int work : {0, 120}) {
        int hops = 20000;
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(cell, 0, 32 * 64);
            cudaEventRecord(e0);
            void* args[] = {&cell, &hops, &lanes, &work};
            cudaLaunchCooperativeKernel((void*)k_chain, dim3(P), dim3(32), args, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("participants %3d lanes %2d work %3d: %.3f us per hop (%s)\n", P, lanes, work, ms * 1e3 / hops, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
