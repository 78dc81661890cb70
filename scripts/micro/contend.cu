// microbenchmark: does polling by the idle warps (mbarrier.try_wait spinning) slow the compute warp's SHFL/LDS chain?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float gs(float x0, float r, float l, float d, float u, float a, float c) {
    float s = __fadd_rn(r, l); s = __fadd_rn(s, d); s = __fadd_rn(s, u);
    return __fmul_rn(__fadd_rn(x0, __fmul_rn(a, s)), c);
}
__device__ __forceinline__ bool try_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
// POLL: 0 idle warps exit, 1 tight try_wait, 2 try_wait + nanosleep(200), 3 tight test_wait, 4 volatile smem flag poll + nanosleep(64)
template <int POLL>
__global__ void k(float *out, long long *cyc, int iters, float a, float c) {
    extern __shared__ float sm[];
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 34 * 128; i += blockDim.x) sm[i] = 0.001f * i;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    }
    __syncthreads();
    if (warp == 0) {
        float cur = lane * 0.5f;
        float *row = sm + (lane + 1) * 128;
        int o = (0 - lane) & 127;
        float r = row[(o + 1) & 127], d = row[128 + o], x0 = row[-128 + o];
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int o1 = (o + 1) & 127;
            float up = __shfl_up_sync(0xffffffffu, cur, 1);
            const float rn = row[(o1 + 1) & 127], dn = row[128 + o1], x0n = row[-128 + o1];
            const float nv = gs(x0, r, cur, d, up, a, c);
            row[(o - 1) & 127] = cur;
            cur = nv; o = o1; r = rn; d = dn; x0 = x0n;
            __syncwarp();
        }
        const long long t1 = clock64();
        out[blockIdx.x * 32 + lane] = cur;
        if (lane == 0) { cyc[blockIdx.x] = t1 - t0; stop = 1; }
    } else if (POLL != 0) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        while (!stop) {
            if (POLL == 1) { if (try_wait(b, 0)) break; }
            else if (POLL == 2) { if (try_wait(b, 0)) break; __nanosleep(200); }
            else if (POLL == 3) { if (test_wait(b, 0)) break; }
            else { __nanosleep(64); }
        }
    }
}
int main() {
    float *out; long long *cyc;
    const int grid = 148 * 5, iters = 50000;
    cudaMalloc(&out, grid * 32 * 4); cudaMalloc(&cyc, grid * 8);
    long long *h = new long long[grid];
#define RUN(P) { cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 44000); \
    k<P><<<grid, 128, 44000>>>(out, cyc, iters, 0.37f, 0.4f); cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost); \
    double s = 0, mx = 0; for (int i = 0; i < grid; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; } \
    printf("poll %d: avg %.1f max %.1f cycles/step  (%s)\n", P, s / grid / iters, mx / iters, cudaGetErrorString(cudaGetLastError())); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(0)
    return 0;
}
