// k_linsolve_rb.cuh -- red-black Gauss-Seidel fast path for lin_solve
// (fluid.rs:301-325 with the sweep order changed): each iteration updates the
// cells with (i+j) even, then the cells with (i+j) odd, then set_boundaries.
// Same formula, same iteration count; results are bit-compared with the red-black
// restatement in the oracle and tolerance-checked against the lexicographic one
// (tests/test_red_black.py).
//
// k_rb_tiled: temporally blocked.  A CTA stages a (TW+2h) x (TH+2h) region (152 x 80 cells) of x, x0 and the
// fix-up codes in shared memory, runs RB_T complete iterations (red, black, boundary fix-up) on
// it and writes the central TW x TH cells: x and x0 cross HBM once per RB_T iterations instead
// of twice per iteration.  Cells near the region edge go stale by 3 cells per iteration (red
// reads black at distance 1, black reads the new red, the fix-up mirrors a neighbour), hence
// h = 3*RB_T; the centre is exact.  Passes ping-pong between two arrays (a tile's halo is
// another tile's output).
#pragma once
#include "eq_common.cuh"

#define RB_T 4
#define RB_H (3 * RB_T)
#define RB_TW 128
#define RB_TH 56
#define RB_RW (RB_TW + 2 * RB_H)      // 152
#define RB_RH (RB_TH + 2 * RB_H)      // 80: two CTAs (109 KB each) fit one SM
#define RB_THREADS 532                 // 7 row phases x 76 column pairs; also 14 rows x 38 quads for staging
#define RB_HW (RB_RW / 2)               // column pairs per row
#define RB_RW4 (RB_RW / 4)              // 4-cell groups per row
#define RB_SMEM_BYTES (RB_RW * RB_RH * 9)

// v1 kernel, kept for the tail (iterations % RB_T) and as the simplest statement of the sweep:
// one launch per colour, thread t of row j owns cell i = 1 + 2t + off.
__global__ void __launch_bounds__(256) k_rb_half(float *__restrict__ x, const float *__restrict__ x0, float a,
                                                 float c_recip, int colour, EqLayout L) {
    const int j = blockIdx.y + max(L.row0, 1);   // owned interior rows
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = 1 + 2 * t + ((colour ^ (j + 1)) & 1);
    if (i > L.N - 2) return;
    const size_t o = (size_t)i + (size_t)j * L.P;
    x[o] = gs_update(x0[o], x[o + 1], x[o - 1], x[o + L.P], x[o - L.P], a, c_recip);
}

// `iters` (<= RB_T) iterations from xin into xout for the tile whose top-left output cell is
// (blockIdx.x*RB_TW, row_lo + blockIdx.y*RB_TH); rows outside [row_lo, row_hi) are not written.
__global__ void __launch_bounds__(RB_THREADS) k_rb_tiled(const float *__restrict__ xin, float *__restrict__ xout,
                                                         const float *__restrict__ x0, const uint8_t *__restrict__ codes,
                                                         const uint8_t *__restrict__ row_fluid,
                                                         const uint8_t *__restrict__ col_fluid, float a, float c_recip,
                                                         int orient, int iters, int row_lo, int row_hi, EqLayout L) {
    EQ_DYN_SMEM(rb_smem);
    float *xs = reinterpret_cast<float *>(rb_smem);
    float *x0s = xs + RB_RW * RB_RH;
    uint8_t *cs = reinterpret_cast<uint8_t *>(x0s + RB_RW * RB_RH);
    const int N = L.N, P = L.P;
    const int gx0 = (int)blockIdx.x * RB_TW - RB_H;              // global coordinates of region cell (0,0)
    const int gy0 = row_lo + (int)blockIdx.y * RB_TH - RB_H;
    // ---- stage the region (cells outside the grid are never touched) ----
    unsigned any_code = 0;
    for (int ly = threadIdx.x / RB_RW4; ly < RB_RH; ly += RB_THREADS / RB_RW4) {
        // RB_RW4 threads per row, each stages 4 consecutive cells (scalar loads: gx0 is not 16-byte aligned)
        const int lx4 = (threadIdx.x % RB_RW4) * 4;
        const int gy = gy0 + ly;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int lx = lx4 + u, gx = gx0 + lx;
            float v = 0.f, v0 = 0.f;
            uint8_t c = EQ_CODE_WALL;
            if (gx >= 0 && gx < N && gy >= 0 && gy < N) {
                const size_t o = (size_t)gx + (size_t)gy * P;
                v = xin[o];
                v0 = x0[o];
                c = codes[o];
            }
            const int p = ly * RB_RW + lx;
            xs[p] = v;
            x0s[p] = v0;
            cs[p] = c;
            any_code |= c & 15u;
        }
    }
    // does this region contain anything the boundary pass has to touch?
    const bool fix_rc = __syncthreads_or((int)any_code) != 0;
    const bool fix_passive = (gx0 <= 0) || (gy0 <= 0) || (gx0 + RB_RW >= N) || (gy0 + RB_RH >= N);
    // thread -> (column pair, row phase): no divisions inside the sweeps
    const int cp = threadIdx.x % RB_HW, ry = threadIdx.x / RB_HW;
    constexpr int RSTEP = RB_THREADS / RB_HW;                       // rows advanced per pass of the row loop
    for (int it = 0; it < iters; ++it) {
        for (int colour = 0; colour < 2; ++colour) {
            // interior cells of the region (all four neighbours staged) that are interior cells of the grid
            if (ry < RSTEP) {
#pragma unroll 4
                for (int ly = 1 + ry; ly <= RB_RH - 2; ly += RSTEP) {
                    const int gy = gy0 + ly;
                    const int lx = 2 * cp + ((colour ^ gy ^ gx0) & 1);
                    const int gx = gx0 + lx;
                    if (lx >= 1 && lx <= RB_RW - 2 && gx >= 1 && gx <= N - 2 && gy >= 1 && gy <= N - 2) {
                        const int o = ly * RB_RW + lx;
                        xs[o] = gs_update(x0s[o], xs[o + 1], xs[o - 1], xs[o + RB_RW], xs[o - RB_RW], a, c_recip);
                    }
                }
            }
            __syncthreads();
        }
        // ---- set_boundaries (fluid.rs:252-272) inside the region ----
        if (orient == EQ_PASSIVE) {
            if (fix_passive) {
                // frame cells copy their interior neighbour (fluid.rs:182-186, conditional per quirk Q6)
                for (int p = threadIdx.x; p < RB_RW * RB_RH; p += RB_THREADS) {
                    const int ly = p / RB_RW, lx = p - ly * RB_RW;
                    const int gx = gx0 + lx, gy = gy0 + ly;
                    if (gx < 0 || gx >= N || gy < 0 || gy >= N) continue;
                    const bool fx = (gx == 0 || gx == N - 1), fy = (gy == 0 || gy == N - 1);
                    if (fx == fy) continue;                       // interior cell or corner
                    if (fy) {
                        if (gx >= 1 && gx <= N - 2 && col_fluid[gx]) {
                            const int src = (gy == 0) ? p + RB_RW : p - RB_RW;
                            if (src >= 0 && src < RB_RW * RB_RH) xs[p] = xs[src];
                        }
                    } else if (gy >= 1 && gy <= N - 2 && row_fluid[gy]) {
                        const int sl = (gx == 0) ? lx + 1 : lx - 1;
                        if (sl >= 0 && sl < RB_RW) xs[p] = xs[ly * RB_RW + sl];
                    }
                }
                __syncthreads();
            }
        } else if (fix_rc) {
            const unsigned shift = (orient == EQ_ADJUST_ROW) ? 0u : 2u;
            for (int p = threadIdx.x; p < RB_RW * RB_RH; p += RB_THREADS) {
                const unsigned code = (cs[p] >> shift) & 3u;
                if (code == 0u || (cs[p] & EQ_CODE_WALL)) continue;
                const int ly = p / RB_RW, lx = p - ly * RB_RW;
                int src;
                if (orient == EQ_ADJUST_ROW) {
                    const int sl = (code == 2u) ? lx + 1 : lx - 1;        // RIGHT : LEFT
                    if (sl < 0 || sl >= RB_RW) continue;
                    src = ly * RB_RW + sl;
                } else {
                    const int sy = (code == 1u) ? ly - 1 : ly + 1;        // UP : DOWN
                    if (sy < 0 || sy >= RB_RH) continue;
                    src = sy * RB_RW + lx;
                }
                xs[p] = -xs[src];                              // sources are wall cells: never a destination
            }
            __syncthreads();
        }
    }
    // ---- write the centre ----
    for (int p = threadIdx.x; p < RB_TW * RB_TH; p += RB_THREADS) {
        const int ty = p / RB_TW, tx = p - ty * RB_TW;
        const int gx = gx0 + RB_H + tx, gy = gy0 + RB_H + ty;
        if (gx < N && gy < N && gy >= row_lo && gy < row_hi) xout[(size_t)gx + (size_t)gy * P] = xs[(ty + RB_H) * RB_RW + tx + RB_H];
    }
}
