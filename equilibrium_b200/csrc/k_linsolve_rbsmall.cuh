// k_linsolve_rbsmall.cuh -- red-black lin_solve (EQ_MODE_RED_BLACK) for grids that fit one SM's shared memory.
//
// The reference's own default scene is 128 x 128 with 100 iterations per solve (configs.rs, BASELINE config 1).  The
// tiled kernels need 25 launches of 4 CTAs for it -- each launch is one CTA's latency, ~50 us -- so a frame took
// 5.5 ms.  Here ONE CTA of 1024 threads keeps the whole grid in shared memory and runs all K iterations in one
// launch: a thread owns a column pair in every RG-th row (x0 and the decoded mirror codes of its cells stay in
// registers), a half-sweep updates the one active cell of each owned pair in place (a colour only reads the other
// colour), set_boundaries runs on the same array (sources are never destinations), three __syncthreads per iteration.
// Same expression tree, colour order and set_boundaries as the other red-black kernels: bit-identical results.
// Limits: N * P <= RBSM_MAX_CELLS (16 owned rows per thread), one GPU.
#pragma once
#include "eq_common.cuh"

#define RBSM_THREADS 1024
#define RBSM_MAXR 16                                  // owned rows per thread
#define RBSM_MAX_CELLS (RBSM_MAXR * RBSM_THREADS * 2)  // 32768: N <= 160

template <int MAXR>   // owned rows per thread: 8 up to 128 x 128 (no spills in 64 registers), 16 up to N * P = 32768
__global__ void __launch_bounds__(RBSM_THREADS, 1) k_rb_small(float *__restrict__ x, const float *__restrict__ x0,
                                                               const uint8_t *__restrict__ codes, float a, float c_recip, int orient,
                                                               int iters, const unsigned *__restrict__ run_if, EqLayout L) {
    EQ_DYN_SMEM(rbsm_smem);
    if (run_if && *run_if == 0u) return;                                   // the a == 0 shortcut was taken (k_a0_check)
    float *sx = reinterpret_cast<float *>(rbsm_smem);
    const int N = L.N, P = L.P, HW = P / 2;
    const int RG = RBSM_THREADS / HW;                                       // row groups; threads >= HW * RG idle in the sweeps
    const int tid = (int)threadIdx.x;
    for (int q = tid; q < N * (P / 4); q += RBSM_THREADS)
        reinterpret_cast<float4 *>(sx)[q] = reinterpret_cast<const float4 *>(x)[q];
    const int tx = tid % HW, ty = tid / HW;
    const bool worker = ty < RG;
    const int i0 = 2 * tx;
    float z[MAXR][2];
    unsigned dc[MAXR];                                                 // mirror direction of my two cells, 4 bits each
#pragma unroll
    for (int m = 0; m < MAXR; ++m) {
        const int j = ty + m * RG;
        z[m][0] = z[m][1] = 0.f;
        dc[m] = 0u;
        if (worker && j < N) {
            const float2 v = *reinterpret_cast<const float2 *>(x0 + (size_t)j * P + i0);
            z[m][0] = v.x;
            z[m][1] = v.y;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const unsigned byte = (i0 + e < N) ? codes[(size_t)j * P + i0 + e] : 0u;
                unsigned d;
                if (orient == EQ_ADJUST_ROW) d = byte & 3u;
                else if (orient == EQ_ADJUST_COLUMN) { const unsigned c = (byte >> 2) & 3u; d = c ? c + 2u : 0u; }
                else d = (byte >> EQ_CODE_PASSIVE_SHIFT) & 7u;
                dc[m] |= d << (4 * e);
            }
        }
    }
    const unsigned sgn = (orient != EQ_PASSIVE) ? 0x80000000u : 0u;
    __syncthreads();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {                                       // red, then black
#pragma unroll
            for (int m = 0; m < MAXR; ++m) {
                const int j = ty + m * RG;
                const int e = (j + c) & 1;                                  // (i + j + c) even, i0 even
                const int i = i0 + e;
                if (worker && j >= 1 && j <= N - 2 && i >= 1 && i <= N - 2) {
                    float *p = sx + j * P + i;
                    *p = gs_update(e ? z[m][1] : z[m][0], p[1], p[-1], p[P], p[-P], a, c_recip);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int m = 0; m < MAXR; ++m) {                               // set_boundaries
            const int j = ty + m * RG;
            if (worker && j < N && dc[m]) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned d = (dc[m] >> (4 * e)) & 15u;
                    if (d) {
                        float *p = sx + j * P + i0 + e;
                        const float src = (d == WF_C_L) ? p[-1] : (d == WF_C_R) ? p[1] : (d == WF_C_U) ? p[-P] : p[P];
                        *p = __uint_as_float(__float_as_uint(src) ^ sgn);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) eq_corners(sx, L);
    __syncthreads();
    for (int q = tid; q < N * (P / 4); q += RBSM_THREADS)
        reinterpret_cast<float4 *>(x)[q] = reinterpret_cast<const float4 *>(sx)[q];
}
