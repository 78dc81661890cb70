// microbenchmark: latency per step of the dependent chains the wavefront solver is made of
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float gs(float x0, float r, float l, float d, float u, float a, float c) {
    float s = __fadd_rn(r, l); s = __fadd_rn(s, d); s = __fadd_rn(s, u);
    return __fmul_rn(__fadd_rn(x0, __fmul_rn(a, s)), c);
}
template <int MODE>
__global__ void k(float *out, long long *cyc, int iters, float a, float c) {
    __shared__ float sm[34 * 128];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 34 * 128; i += blockDim.x) sm[i] = 0.001f * i;
    __syncthreads();
    float cur = lane * 0.5f;
    const long long t0 = clock64();
    if (MODE == 0) {           // shuffle only
        for (int i = 0; i < iters; ++i) cur = __shfl_up_sync(0xffffffffu, cur, 1) + 1.0f;
    } else if (MODE == 1) {    // shuffle + gs chain, operands in registers
        for (int i = 0; i < iters; ++i) {
            float up = __shfl_up_sync(0xffffffffu, cur, 1);
            cur = gs(0.3f, 0.1f, cur, 0.2f, up, a, c);
        }
    } else if (MODE == 2) {    // + 3 LDS + 1 STS per step (ring addressing), syncwarp
        float *row = sm + (lane + 1) * 128;
        for (int i = 0; i < iters; ++i) {
            const int o = (i - lane) & 127;
            float up = __shfl_up_sync(0xffffffffu, cur, 1);
            const float r = row[(o + 1) & 127], d = row[128 + o], x0 = row[-128 + o];
            const float nv = gs(x0, r, cur, d, up, a, c);
            row[(o - 1) & 127] = cur;
            cur = nv;
            __syncwarp();
        }
    } else if (MODE == 3) {    // like 2 but operands of the next step prefetched
        float *row = sm + (lane + 1) * 128;
        int o = (0 - lane) & 127;
        float r = row[(o + 1) & 127], d = row[128 + o], x0 = row[-128 + o];
        for (int i = 0; i < iters; ++i) {
            const int o1 = (o + 1) & 127;
            float up = __shfl_up_sync(0xffffffffu, cur, 1);
            const float rn = row[(o1 + 1) & 127], dn = row[128 + o1], x0n = row[-128 + o1];
            const float nv = gs(x0, r, cur, d, up, a, c);
            row[(o - 1) & 127] = cur;
            cur = nv; o = o1; r = rn; d = dn; x0 = x0n;
            __syncwarp();
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = cur;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int iters = 100000;
#define RUN(M, W) k<M><<<1, 32 * W>>>(out, cyc, iters, 0.37f, 0.4f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("mode %d warps %d: %.1f cycles/step\n", M, W, (double)h / iters);
    RUN(0, 1) RUN(1, 1) RUN(2, 1) RUN(3, 1) RUN(1, 4) RUN(3, 4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
