// microbenchmark for the register-resident wavefront step (k_linsolve_wf design study):
// one compute warp per CTA runs T sub-steps per step; sub-step t+1 takes its "old" operands from sub-step t's
// registers by shuffle (rows move up SH lanes per sub-step), x0 from a de-skewed shared tile ([reg + imm] addresses),
// band edges through one predicated LDS.128 (lanes < 2) and one predicated STS.64 (lanes >= 30).
// usage: wfstep [ctas_per_sm]     prints cycles per step for several (T, SH)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float gs(float x0, float r, float l, float d, float u, float a, float c) {
    float s = __fadd_rn(r, l); s = __fadd_rn(s, d); s = __fadd_rn(s, u);
    return __fmul_rn(__fadd_rn(x0, __fmul_rn(a, s)), c);
}

#define RING 96
#define STRIDE 113   // floats per tile row (odd: the 32 lanes of a step hit 32 banks)

template <int T, int SH, bool EDGES>
__global__ void __launch_bounds__(64) k(float *out, long long *cyc, int macros, float a, float c) {
    extern __shared__ __align__(16) float sm[];
    float *xin = sm;                               // 33 rows
    float *x0t = xin + 33 * STRIDE;                // 32 + 2(T-1) rows
    float *outt = x0t + (32 + 2 * (T - 1)) * STRIDE;   // 32 rows
    float4 *ein = reinterpret_cast<float4 *>(sm + (((33 + 32 + 2 * (T - 1) + 32) * STRIDE + 3) & ~3));   // T x (RING+16) float4
    float2 *eout = reinterpret_cast<float2 *>(ein + T * (RING + 16));                                    // T x 2 x (RING+16) float2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 *ein_l = ein;             // every lane reads the same entry (broadcast), lanes 0/1 use it
    float2 *eout_l = eout + (lane & 1) * (RING + 16);
    const int total = 33 * STRIDE + (32 + 2 * (T - 1)) * STRIDE + 32 * STRIDE + 8 + T * (RING + 16) * 4 + T * 2 * (RING + 16) * 2;
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 1e-3f * (float)(i % 977);
    __syncthreads();
    if (warp != 0) return;
    float cur[T], fh[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { cur[t] = 0.25f * lane + t; fh[t] = 0.5f * lane - t; }
    const float *xr = xin + lane * STRIDE, *xd = xin + (lane + 1) * STRIDE;
    const float *x0r[T];
#pragma unroll
    for (int t = 0; t < T; ++t) x0r[t] = x0t + (2 * (T - 1 - t) + lane) * STRIDE;
    float *orow = outt + lane * STRIDE;
    const long long t0 = clock64();
    int off = 0;
    for (int m = 0; m < macros; ++m) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float up[T], rgt[T], dwn[T], x0v[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                up[t] = __shfl_up_sync(0xffffffffu, cur[t], 1);
                if (t == 0) {
                    rgt[t] = xr[off + i];
                    dwn[t] = xd[off + i];
                } else {
                    rgt[t] = __shfl_up_sync(0xffffffffu, fh[t - 1], SH);
                    dwn[t] = (SH == 2) ? __shfl_up_sync(0xffffffffu, fh[t - 1], 1) : fh[t - 1];
                }
                x0v[t] = x0r[t][off + i + 3 * (T - 1 - t)];
                if (EDGES) {
                    const float4 e = ein_l[t * (RING + 16) + off + i];
                    up[t] = lane == 0 ? e.z : up[t];
                    if (t) {
                        rgt[t] = lane == 0 ? e.x : rgt[t];
                        rgt[t] = lane == 1 ? e.w : rgt[t];
                        dwn[t] = lane == 0 ? e.y : dwn[t];
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const float nv = gs(x0v[t], rgt[t], cur[t], dwn[t], up[t], a, c);
                fh[t] = cur[t];
                cur[t] = nv;
                if (EDGES) {
                    if (lane >= 30) eout_l[t * 2 * (RING + 16) + off + i] = make_float2(fh[t], nv);
                }
                if (t == T - 1) orow[off + i] = fh[t];
            }
        }
        off += 16;
        if (off >= RING) off -= RING;
        __syncwarp();
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) s += cur[t] + fh[t];
    out[blockIdx.x * 32 + lane] = s;
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T, int SH, bool EDGES>
static void run(int per_sm, float *out, long long *cyc) {
    const int macros = 4000;
    const size_t smem = (size_t)(33 + 32 + 2 * (T - 1) + 32) * STRIDE * 4 + 64 + (size_t)T * (RING + 16) * 16 + (size_t)T * 2 * (RING + 16) * 8;
    cudaFuncSetAttribute(k<T, SH, EDGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<T, SH, EDGES>, 64, smem);
    const int grid = 148 * per_sm;
    k<T, SH, EDGES><<<grid, 64, smem>>>(out, cyc, macros, 0.37f, 0.4f);
    static long long h[148 * 16];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    const double per_step = avg / (macros * 16.0);
    printf("T=%d SH=%d edges=%d ctas/SM=%d (occ %d, smem %zu): %.1f cycles/step  -> %.2f cells/cycle/SM  (%s)\n", T, SH, (int)EDGES,
           per_sm, occ, smem, per_step, per_sm * 32.0 * T / per_step, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 16 * 32 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    for (int per_sm = 1; per_sm <= 4; ++per_sm) {
        run<2, 2, true>(per_sm, out, cyc);
        run<4, 2, true>(per_sm, out, cyc);
        run<4, 1, true>(per_sm, out, cyc);
        run<4, 2, false>(per_sm, out, cyc);
        run<8, 2, true>(per_sm, out, cyc);
        run<8, 1, true>(per_sm, out, cyc);
    }
    return 0;
}
