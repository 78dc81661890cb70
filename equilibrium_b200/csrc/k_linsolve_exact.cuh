// k_linsolve_exact.cuh -- bit-exact lexicographic Gauss-Seidel (fluid.rs:301-325)
// as a space-time wavefront of warp-sized jobs.
//
// Reference semantics being reproduced:
//   repeat K times { for j in 1..N-1 { for i in 1..N-1 { x[i,j] = GS(...) in place } }
//                    set_boundaries(orientation, x) }
// Cell (i,j) of iteration k reads (i-1,j),(i,j-1) of iteration k *before* that
// iteration's set_boundaries ("raw" values R_k) and (i+1,j),(i,j+1) of iteration
// k-1 *after* its set_boundaries ("fixed" values F_{k-1}).
//
// Decomposition (DESIGN.md "Exact lin_solve"):
//   job (b,k)  = rows j0..j0+31 (j0 = 1+32b) of iteration k, run by ONE warp,
//                lane r = row j0+r, lane r trails lane r-1 by one column, so the
//                warp walks an anti-diagonal along the row; R_k(i,j-1) arrives by
//                __shfl_up, R_k(i-1,j) is the lane's own previous result.
//   job (b,k) needs (b-1,k)   : R_k of row j0-1 (a "raw" side stream, because
//                               global x only ever holds fixed values), and
//             and (b+1,k-1)   : F_{k-1} of rows j0..j0+32.
//   Jobs are handed out by an atomic ticket in wavefront order w = b + 2k, so a
//   job only ever waits for lower tickets (already running or done): no
//   co-residency requirement, no deadlock.  Progress is published per 32-column
//   chunk with release/acquire flags.
//   The boundary fix-up of set_boundaries is fused: a cell is written once, one
//   step after it was computed, with its fixed value (left/right/up sources are
//   in registers, the down source comes by __shfl_down; across a band edge the
//   lower band patches the one cell above it).
//
// HBM traffic per cell-iteration: read x, x0, write x (12 B) + 1 B code in the
// AdjustRow/AdjustColumn orientations + 8/32 B for the raw stream.
#pragma once
#include "eq_common.cuh"

#define LSX_SLOTS 4
#define LSX_XROWS 34   // rows j0-1 .. j0+32
#define LSX_CROWS 33   // code rows j0-1 .. j0+31
#define LSX_SPIN_LIMIT (1u << 22)

struct LsxProblem {
    float *x;            // in/out, in place
    const float *x0;
    float *raw;          // [NB][P] raw stream: R_k of the last row of band b-1, read by band b
    unsigned *progress;  // [K][NB] chunks completed
    float a, c_recip;
    int orient;          // EqOrientation
};

struct LsxParams {
    LsxProblem prob[2];
    int nprob;
    const uint8_t *codes;      // per-cell fix-up codes, pitch P
    const uint8_t *row_fluid;  // [N] row j has a NoWall cell   (quirk Q6)
    const uint8_t *col_fluid;  // [N] column i has a NoWall cell
    const uint32_t *jobs;      // [K*NB] (k << 16 | b) in wavefront order
    int njobs;                 // K*NB (per problem)
    int N, P, K, NB, NC;
    unsigned *ticket;
    int *error;
};

struct LsxSmem {
    float xs[LSX_SLOTS][LSX_XROWS][32];
    float x0s[LSX_SLOTS][32][32];
    uint8_t cs[LSX_SLOTS][LSX_CROWS][32];
    float rawbuf[64];
};

__device__ __forceinline__ bool lsx_wait_ge(const unsigned *flag, unsigned need, int *error) {
    unsigned spins = 0;
    while (ld_acquire_u32(flag) < need) {
        __nanosleep(40);
        if ((++spins & 1023u) == 0) {
            if (spins >= LSX_SPIN_LIMIT) {
                *error = 1;
                return false;
            }
            if (ld_volatile_s32(error) != 0) return false;
        }
    }
    return true;
}

template <int ORIENT>
__device__ __forceinline__ bool lsx_run_job(const LsxParams &p, const LsxProblem &pr, LsxSmem &sm,
                                            const int b, const int k) {
    const int lane = threadIdx.x;
    const int N = p.N, P = p.P, NC = p.NC, NB = p.NB;
    const int j0 = 1 + 32 * b;
    const int j = j0 + lane;
    const int tr = lane + 1;
    const bool in_row = (j <= N - 2);
    const bool last_band = (b == NB - 1);
    const float a = pr.a, c_recip = pr.c_recip;
    float *__restrict__ x = pr.x;
    const float *__restrict__ x0 = pr.x0;
    const unsigned *flag_prev_iter = (k > 0) ? pr.progress + (size_t)(k - 1) * NB + min(b + 1, NB - 1) : nullptr;
    const unsigned *flag_band_above = (b > 0) ? pr.progress + (size_t)k * NB + (b - 1) : nullptr;
    unsigned *my_flag = pr.progress + (size_t)k * NB + b;
    const float *top_src = (b > 0) ? pr.raw + (size_t)b * P : x;  // row j0-1: raw stream or frame row 0
    float *raw_out = (b + 1 < NB) ? pr.raw + (size_t)(b + 1) * P : nullptr;
    const bool row_has_fluid = (ORIENT == EQ_PASSIVE && in_row) ? (p.row_fluid[j] != 0) : false;

    auto load_chunk = [&](int q) -> bool {
        if (q < NC) {
            if (flag_prev_iter && !lsx_wait_ge(flag_prev_iter, (unsigned)q + 1u, p.error)) return false;
            if (flag_band_above && !lsx_wait_ge(flag_band_above, (unsigned)q + 1u, p.error)) return false;
            const int slot = q & (LSX_SLOTS - 1);
            const int col0 = 32 * q;
            {   // x rows j0-1 .. j0+32 : 8 lanes x 16 B per row, 4 rows per pass
                const int sub = lane & 7, rr = lane >> 3;
#pragma unroll
                for (int g = 0; g < 9; ++g) {
                    const int t = 4 * g + rr;
                    if (t < LSX_XROWS) {
                        const float *src = (t == 0) ? top_src + col0 + 4 * sub
                                                    : x + (size_t)(j0 - 1 + t) * P + col0 + 4 * sub;
                        cp_async_16(&sm.xs[slot][t][4 * sub], src);
                    }
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int t = 4 * g + rr;
                    cp_async_16(&sm.x0s[slot][t][4 * sub], x0 + (size_t)(j0 + t) * P + col0 + 4 * sub);
                }
            }
            if (ORIENT != EQ_PASSIVE) {   // codes rows j0-1 .. j0+31 : 2 lanes x 16 B per row
                const int sub = lane & 1, rr = lane >> 1;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const int t = 16 * g + rr;
                    if (t < LSX_CROWS)
                        cp_async_16(&sm.cs[slot][t][16 * sub], p.codes + (size_t)(j0 - 1 + t) * P + col0 + 16 * sub);
                }
            }
        }
        cp_async_commit();
        return true;
    };

    auto store_chunk = [&](int q) {
        const int slot = q & (LSX_SLOTS - 1);
        const int col0 = 32 * q;
        const int sub = lane & 7, rr = lane >> 3;
        // band rows (tile rows 1..32); the bottom frame row N-1 travels with the last band
        const int t_hi = last_band ? 33 : 32;
        const int t_lo = (ORIENT == EQ_PASSIVE && b == 0) ? 0 : 1;  // Passive rewrites frame row 0
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            const int t = 4 * g + rr;
            const int row = j0 - 1 + t;
            if (t >= t_lo && t <= t_hi && row <= N - 1) {
                const float4 v = *reinterpret_cast<const float4 *>(&sm.xs[slot][t][4 * sub]);
                *reinterpret_cast<float4 *>(x + (size_t)row * P + col0 + 4 * sub) = v;
            }
        }
        if (raw_out) raw_out[col0 + lane] = sm.rawbuf[(col0 + lane) & 63];
    };

    auto publish = [&](unsigned chunks_done) {
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(my_flag, chunks_done);
    };

#define XS(t, col) sm.xs[((col) >> 5) & (LSX_SLOTS - 1)][(t)][(col) & 31]

    // prologue: chunks 0 and 1
    if (!load_chunk(0)) return false;
    if (!load_chunk(1)) return false;

    float cur = 0.f, prev2 = 0.f, prev_up = 0.f;
    const int S = N + 31;                    // steps 0 .. N+30 (lane r computes column s-r)
    const int M = (S + 31) >> 5;
    int stored = 0;                          // chunks stored so far

    for (int m = 0; m < M; ++m) {
        if (!load_chunk(m + 2)) return false;
        cp_async_wait<1>();                  // chunk m+1 (and older) has landed
        __syncwarp();

        const int s_end = min(32 * m + 32, S);
        for (int s = 32 * m; s < s_end; ++s) {
            const int c = s - lane;          // column this lane computes now (0 = left frame cell)
            const float up = __shfl_up_sync(0xffffffffu, cur, 1);
            float newv = cur;
            float top = up;
            if (c >= 0 && c <= N - 1) {
                if (in_row && c >= 1 && c <= N - 2) {
                    const float right = XS(tr, c + 1);
                    const float down = XS(tr + 1, c);
                    if (lane == 0) top = XS(0, c);
                    const float x0v = sm.x0s[(c >> 5) & (LSX_SLOTS - 1)][lane][c & 31];
                    newv = gs_update(x0v, right, cur, down, top, a, c_recip);
                } else if (in_row || ORIENT == EQ_ADJUST_COLUMN) {
                    newv = XS(tr, c);        // frame column / frame row N-1: pass through
                }
            }
            float dn = 0.f;
            if (ORIENT == EQ_ADJUST_COLUMN) {
                dn = __shfl_down_sync(0xffffffffu, newv, 1);
            }
            const int cf = c - 1;            // column finalised now
            if (in_row && cf >= 1 && cf <= N - 2) {
                float F = cur;               // R_k(cf, j)
                if (ORIENT == EQ_ADJUST_ROW) {
                    const unsigned code = sm.cs[(cf >> 5) & (LSX_SLOTS - 1)][tr][cf & 31] & 3u;
                    if (code == EQ_CODE_ROW_RIGHT) F = -newv;
                    else if (code == EQ_CODE_ROW_LEFT) F = -prev2;
                } else if (ORIENT == EQ_ADJUST_COLUMN) {
                    const unsigned code = sm.cs[(cf >> 5) & (LSX_SLOTS - 1)][tr][cf & 31] & 12u;
                    if (code == EQ_CODE_COL_UP) F = -prev_up;
                    else if (code == EQ_CODE_COL_DOWN) {
                        if (lane < 31) F = -dn;
                        else if (last_band) F = -XS(33, cf);
                        // else: the band below patches this cell (see below)
                    }
                }
                XS(tr, cf) = F;
                if (ORIENT == EQ_PASSIVE) {   // fluid.rs:179-187, conditional per quirk Q6
                    if (row_has_fluid) {
                        if (cf == 1) XS(tr, 0) = cur;
                        if (cf == N - 2) XS(tr, N - 1) = cur;
                    }
                    if ((j == 1 || j == N - 2) && p.col_fluid[cf]) {
                        if (j == 1) XS(0, cf) = cur;
                        if (j == N - 2) XS(tr + 1, cf) = cur;
                    }
                }
            }
            if (in_row && c >= 1 && c <= N - 2) {
                if (ORIENT == EQ_ADJUST_COLUMN && lane == 0 && b > 0) {
                    // cell (c, j0-1) of the band above takes -R_k(c, j0) when its code says DOWN
                    const unsigned code0 = sm.cs[(c >> 5) & (LSX_SLOTS - 1)][0][c & 31] & 12u;
                    if (code0 == EQ_CODE_COL_DOWN) x[(size_t)(j0 - 1) * P + c] = -newv;
                }
                if (lane == 31) sm.rawbuf[c & 63] = newv;
            }
            prev2 = cur;
            prev_up = top;
            cur = newv;
            __syncwarp();
        }

        if (m >= 1 && m - 1 < NC) {          // columns < 32m are final for every lane
            store_chunk(m - 1);
            stored = m;
            publish((unsigned)m);
        }
    }
    for (int q = stored; q < NC; ++q) store_chunk(q);
    publish((unsigned)NC);
    cp_async_wait<0>();
    __syncwarp();
#undef XS
    return true;
}

// One warp per CTA; persistent CTAs pull jobs from the ticket counter.
__global__ void __launch_bounds__(32) k_linsolve_exact(const LsxParams p) {
    EQ_DYN_SMEM(lsx_smem_raw);
    LsxSmem &sm = *reinterpret_cast<LsxSmem *>(lsx_smem_raw);
    const int total = p.njobs * p.nprob;
    if (threadIdx.x == 0) sm.rawbuf[0] = 0.f;   // column 0 of the raw stream is never produced
    for (;;) {
        if (ld_volatile_s32(p.error) != 0) break;
        unsigned t = 0;
        if (threadIdx.x == 0) t = atomicAdd(p.ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((int)t >= total) break;
        const int pi = (int)(t % (unsigned)p.nprob);
        const uint32_t jb = p.jobs[t / (unsigned)p.nprob];
        const int k = (int)(jb >> 16), b = (int)(jb & 0xffffu);
        const LsxProblem &pr = p.prob[pi];
        bool ok;
        if (pr.orient == EQ_ADJUST_ROW) ok = lsx_run_job<EQ_ADJUST_ROW>(p, pr, sm, b, k);
        else if (pr.orient == EQ_ADJUST_COLUMN) ok = lsx_run_job<EQ_ADJUST_COLUMN>(p, pr, sm, b, k);
        else ok = lsx_run_job<EQ_PASSIVE>(p, pr, sm, b, k);
        if (!ok) break;
    }
}
