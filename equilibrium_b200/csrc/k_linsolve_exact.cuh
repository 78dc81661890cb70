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
//                lane r = row j0+r, lane r trails lane r-1 by one column (at step s
//                lane r computes column s-r), so the warp walks an anti-diagonal
//                along the row; R_k(i,j-1) arrives by __shfl_up, R_k(i-1,j) is the
//                lane's own previous result.
//   job (b,k) needs (b-1,k)   : R_k of row j0-1 (a "raw" side stream, because
//                               global x only ever holds fixed values), and
//                   (b+1,k-1) : F_{k-1} of rows j0..j0+32.
//   Jobs are handed out by an atomic ticket in wavefront order w = b + 2k, so a
//   job only ever waits for lower tickets (already running or done): no
//   co-residency requirement, no deadlock.  Progress is published per 32-column
//   chunk with release/acquire flags.
//   The boundary fix-up of set_boundaries is fused: a cell is written once, one
//   step after it was computed, with its fixed value (left/right/up sources are
//   in registers, the down source comes by __shfl_down; across a band edge the
//   lower band patches the one cell above it).
//
// Shared memory per warp (38 272 B): a 4-chunk ring (128 columns) of
//   x   rows j0-1 .. j0+32   (34 x 512 B)     addr = tr*512 + (c & 127)*4
//   x0  rows j0   .. j0+31   (32 x 512 B)
//   fix-up codes rows j0-1 .. j0+31 (33 x 128 B)
// A lane reads column c of its own row while its neighbours read c+1 / c-1: the
// row pitch is a multiple of 32 banks, so the diagonal access is conflict-free.
// All shared accesses use explicit 32-bit shared addresses (no generic->shared
// conversion in the loop).
//
// Warp specialisation: a CTA is three warps working on one job -- a LOADER (polls the
// dependency flags with ld.acquire, stages chunks with cp.async and signals an mbarrier), the
// COMPUTE warp (never touches global memory in its loop) and a STORER (writes finished chunks
// back, then publishes progress with st.release).  The release fence (MEMBAR, ~2 us while the
// chunk's stores drain) therefore stalls only the storer; ncu showed it costing 29 % of all
// stall samples when the compute warp issued it itself (profiles/r01_linsolve_notes.md).
//
// Macro step m = steps 32m..32m+31 touches chunks m-1, m, m+1.  When every lane is
// on an interior column and the (band, chunk) summary says no cell of those chunks
// has a fix-up code, the macro step runs a branch-free fast loop (3 LDS, 1 SHFL,
// 6 FP ops, 1 STS per step); otherwise a general loop handles frame columns,
// fix-ups and the Passive frame copies.
//
// HBM traffic per cell-iteration: read x, x0, write x (12 B) + 1 B code only for
// chunks that contain fix-ups + 8/32 B for the raw stream.
#pragma once
#include "eq_common.cuh"

#define LSX_XROWS 34   // rows j0-1 .. j0+32
#define LSX_CROWS 33   // code rows j0-1 .. j0+31
#define LSX_XS_OFF 0u
#define LSX_X0_OFF (LSX_XROWS * 512u)
#define LSX_CS_OFF (LSX_X0_OFF + 32u * 512u)
#define LSX_RAW_OFF (LSX_CS_OFF + LSX_CROWS * 128u)    // 128-float ring of the raw stream (lane 31's results)
#define LSX_BAR_OFF (LSX_RAW_OFF + 128u * 4u)          // mbarriers: full[4], done[4], free[4], 16 B each
#define LSX_MISC_OFF (LSX_BAR_OFF + 12u * 16u)         // [0] ticket broadcast
#define LSX_SMEM_BYTES (LSX_MISC_OFF + 16u)
#define LSX_THREADS 96
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
    const uint8_t *codes;        // per-cell fix-up codes, pitch P
    const uint8_t *chunk_flags;  // [2][NB][NC]: (band, chunk) holds an AdjustRow ([0]) / AdjustColumn ([1]) code
    const uint8_t *row_fluid;    // [N] row j has a NoWall cell   (quirk Q6)
    const uint8_t *col_fluid;    // [N] column i has a NoWall cell
    const uint32_t *jobs;        // [K*NB] (k << 16 | b) in wavefront order
    int njobs;                   // K*NB (per problem)
    int N, P, K, NB, NC;
    unsigned *ticket;
    int *error;
};

// ---- waiting primitives: lane 0 waits, the result is broadcast; every loop can be aborted ----
// Dependency flags of other jobs (global memory, acquire).
__device__ __forceinline__ bool lsx_wait_flags(const unsigned *f1, unsigned n1, const unsigned *f2, unsigned n2,
                                               int *error, int lane) {
    int ok = 1;
    if (lane == 0) {
        unsigned spins = 0;
        while ((f1 && ld_acquire_u32(f1) < n1) || (f2 && ld_acquire_u32(f2) < n2)) {
            __nanosleep(32);
            if ((++spins & 1023u) == 0) {
                if (spins >= LSX_SPIN_LIMIT) {
                    *error = 1;
                    ok = 0;
                    break;
                }
                if (ld_volatile_s32(error) != 0) {
                    ok = 0;
                    break;
                }
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok != 0;
}
// An mbarrier of this CTA.
__device__ __forceinline__ bool lsx_wait_bar(uint32_t bar, uint32_t parity, int *error, int lane) {
    int ok = 1;
    if (lane == 0) {
        unsigned spins = 0;
        while (!mbar_try_wait(bar, parity)) {
            if ((++spins & 255u) == 0) {
                if (spins >= LSX_SPIN_LIMIT) {
                    *error = 2;
                    ok = 0;
                    break;
                }
                if (ld_volatile_s32(error) != 0) {
                    ok = 0;
                    break;
                }
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok != 0;
}

// Everything the three roles of a job share.
template <int ORIENT>
struct LsxJob {
    const LsxParams &p;
    const LsxProblem &pr;
    uint32_t sbase;
    int b, k, lane;
    int N, P, NC, NB, j0, M;
    bool last_band, full_band;
    const uint8_t *cflags;

    __device__ __forceinline__ LsxJob(const LsxParams &p_, const LsxProblem &pr_, uint32_t sbase_, int b_, int k_,
                                      int lane_)
        : p(p_), pr(pr_), sbase(sbase_), b(b_), k(k_), lane(lane_) {
        N = p.N; P = p.P; NC = p.NC; NB = p.NB;
        j0 = 1 + 32 * b;
        M = (N + 31 + 31) >> 5;                       // steps 0 .. N+30 in macro steps of 32
        last_band = (b == NB - 1);
        full_band = (j0 + 31 <= N - 2);
        cflags = p.chunk_flags + (ORIENT == EQ_ADJUST_COLUMN ? (size_t)NB * NC : 0) + (size_t)b * NC;
    }
    __device__ __forceinline__ uint32_t bar_full(int q) const { return sbase + LSX_BAR_OFF + (uint32_t)(q & 3) * 16u; }
    __device__ __forceinline__ uint32_t bar_done(int q) const { return sbase + LSX_BAR_OFF + 64u + (uint32_t)(q & 3) * 16u; }
    __device__ __forceinline__ uint32_t bar_free(int q) const { return sbase + LSX_BAR_OFF + 128u + (uint32_t)(q & 3) * 16u; }
    // How macro step m (steps 32m..32m+31, columns 32m-31..32m+31) is executed:
    //   EDGE   some lane is on/next to a frame column: fully general loop (first two and last few macro steps)
    //   CODED  interior columns, but chunks m-1/m hold fix-up codes of this orientation, or (Passive)
    //          the band owns a frame row: branch-light loop that reads one code byte per step
    //   FAST   interior columns, nothing to fix up: 3 LDS + 1 SHFL + 6 FP + 1 STS per step
    // Every job runs at the pace of the slowest job it depends on, so CODED must stay close to FAST.
    enum { MODE_FAST = 0, MODE_CODED = 1, MODE_EDGE = 2 };
    __device__ __forceinline__ int macro_mode(int m) const {
        if (m < 2 || 32 * m + 31 > N - 2) return MODE_EDGE;
        if (ORIENT == EQ_PASSIVE) return (b == 0 || last_band) ? MODE_CODED : MODE_FAST;
        return (cflags[m - 1] | cflags[m]) ? MODE_CODED : MODE_FAST;
    }
    // the code tile of chunk q is read by macro steps q and q+1
    __device__ __forceinline__ bool need_codes(int q) const {
        if (ORIENT == EQ_PASSIVE) return b == 0 || last_band;      // row 0 of the tile carries col_fluid
        return macro_mode(q) != MODE_FAST || macro_mode(q + 1) != MODE_FAST;
    }

    // ------------------------------------------------------------------ LOADER warp
    __device__ __forceinline__ bool run_loader() const {
        const float *__restrict__ x = pr.x;
        const float *__restrict__ x0 = pr.x0;
        const unsigned *flag_prev_iter = (k > 0) ? pr.progress + (size_t)(k - 1) * NB + min(b + 1, NB - 1) : nullptr;
        const unsigned *flag_band_above = (b > 0) ? pr.progress + (size_t)k * NB + (b - 1) : nullptr;
        const float *top_src = (b > 0) ? pr.raw + (size_t)b * P : x;  // row j0-1: raw stream or frame row 0
        for (int q = 0; q < NC; ++q) {
            // the ring slot must have been written back (chunk q-4) ...
            if (q >= 4 && !lsx_wait_bar(bar_free(q), (uint32_t)(((q >> 2) - 1) & 1), p.error, lane)) return false;
            // ... and the producers of this chunk must have published it
            if (!lsx_wait_flags(flag_prev_iter, (unsigned)q + 1u, flag_band_above, (unsigned)q + 1u, p.error, lane))
                return false;
            const uint32_t slot = (uint32_t)(q & 3) * 128u;   // byte offset of the chunk inside a 512 B row
            const int col0 = 32 * q;
            {   // x rows j0-1 .. j0+32 : 8 lanes x 16 B per row, 4 rows per pass
                const int sub = lane & 7, rr = lane >> 3;
#pragma unroll
                for (int g = 0; g < 9; ++g) {
                    const int t = 4 * g + rr;
                    if (t < LSX_XROWS) {
                        const float *src = (t == 0) ? top_src + col0 + 4 * sub
                                                    : x + (size_t)(j0 - 1 + t) * P + col0 + 4 * sub;
                        cp_async_16s(sbase + LSX_XS_OFF + (uint32_t)t * 512u + slot + 16u * sub, src);
                    }
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int t = 4 * g + rr;
                    cp_async_16s(sbase + LSX_X0_OFF + (uint32_t)t * 512u + slot + 16u * sub,
                                 x0 + (size_t)(j0 + t) * P + col0 + 4 * sub);
                }
            }
            // codes rows j0-1 .. j0+31 (2 lanes x 16 B per row), skipped when both macro steps
            // that touch this chunk take the fast loop.  Passive has no codes; a band that owns a
            // frame row stages col_fluid (quirk Q6) in row 0 of the code tile instead.
            if (need_codes(q)) {
                if (ORIENT == EQ_PASSIVE) {
                    if (lane < 2)
                        cp_async_16s(sbase + LSX_CS_OFF + (uint32_t)(q & 3) * 32u + 16u * lane,
                                     p.col_fluid + col0 + 16 * lane);
                } else {
                    const int sub = lane & 1, rr = lane >> 1;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const int t = 16 * g + rr;
                        if (t < LSX_CROWS)
                            cp_async_16s(sbase + LSX_CS_OFF + (uint32_t)t * 128u + (uint32_t)(q & 3) * 32u + 16u * sub,
                                         p.codes + (size_t)(j0 - 1 + t) * P + col0 + 16 * sub);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(bar_full(q));     // fires when this lane's copies have landed
        }
        return true;
    }

    // ------------------------------------------------------------------ STORER warp
    __device__ __forceinline__ bool run_storer() const {
        float *__restrict__ x = pr.x;
        unsigned *my_flag = pr.progress + (size_t)k * NB + b;
        float *raw_out = (b + 1 < NB) ? pr.raw + (size_t)(b + 1) * P : nullptr;
        const uint32_t raw_s = sbase + LSX_RAW_OFF;
        const int sub = lane & 7, rr = lane >> 3;
        // band rows (tile rows 1..32); the bottom frame row N-1 travels with the last band
        const int t_hi = last_band ? 33 : 32;
        const int t_lo = (ORIENT == EQ_PASSIVE && b == 0) ? 0 : 1;  // Passive rewrites frame row 0
        for (int q = 0; q < NC; ++q) {
            if (!lsx_wait_bar(bar_done(q), (uint32_t)((q >> 2) & 1), p.error, lane)) return false;
            const uint32_t slot = (uint32_t)(q & 3) * 128u;
            const int col0 = 32 * q;
#pragma unroll
            for (int g = 0; g < 9; ++g) {
                const int t = 4 * g + rr;
                const int row = j0 - 1 + t;
                if (t >= t_lo && t <= t_hi && row <= N - 1) {
                    const float4 v = lds_f32x4(sbase + LSX_XS_OFF + (uint32_t)t * 512u + slot + 16u * sub);
                    *reinterpret_cast<float4 *>(x + (size_t)row * P + col0 + 4 * sub) = v;
                }
            }
            if (raw_out) raw_out[col0 + lane] = lds_f32(raw_s + (uint32_t)((col0 + lane) & 127) * 4u);
            __syncwarp();                             // every lane's loads and stores are issued ...
            if (lane == 0) {
                mbar_arrive(bar_free(q));             // ... the ring slot may be refilled,
                st_release_u32(my_flag, (unsigned)q + 1u);   // and the release makes the stores visible GPU-wide
            }
        }
        return true;
    }

    // ------------------------------------------------------------------ COMPUTE warp
    __device__ __forceinline__ bool run_compute() const {
        const int j = j0 + lane;
        const int tr = lane + 1;
        const bool in_row = (j <= N - 2);
        const float a = pr.a, c_recip = pr.c_recip;
        float *__restrict__ x = pr.x;
        const bool row_has_fluid = (ORIENT == EQ_PASSIVE && in_row) ? (p.row_fluid[j] != 0) : false;
        // shared addresses of this lane's rows
        const uint32_t xs_row = sbase + LSX_XS_OFF + (uint32_t)tr * 512u;     // own row
        const uint32_t xs_top = sbase + LSX_XS_OFF;                            // row j0-1
        const uint32_t x0_row = sbase + LSX_X0_OFF + (uint32_t)lane * 512u;
        const uint32_t cs_row = sbase + LSX_CS_OFF + (uint32_t)tr * 128u;
        const uint32_t cs_top = sbase + LSX_CS_OFF;
        const uint32_t raw_s = sbase + LSX_RAW_OFF;
        const int S = N + 31;                    // steps 0 .. N+30

        if (!lsx_wait_bar(bar_full(0), 0u, p.error, lane)) return false;
        float cur = 0.f, prev2 = 0.f, prev_up = 0.f;

        for (int m = 0; m < M; ++m) {
            // macro step m reads chunks m-1, m, m+1
            if (m + 1 < NC && !lsx_wait_bar(bar_full(m + 1), (uint32_t)(((m + 1) >> 2) & 1), p.error, lane)) return false;

            const int mode = macro_mode(m);
            if (mode != MODE_EDGE) {
                // ---- interior columns: every lane computes and finalises an interior cell ---------
                uint32_t o = ((uint32_t)(32 * m - lane) & 127u) << 2;      // byte offset of column c
                uint32_t om1 = (o - 4u) & 508u;                            // column c-1
                if (mode == MODE_FAST) {
#pragma unroll 4
                    for (int t = 0; t < 32; ++t) {
                        const uint32_t o1 = (o + 4u) & 508u;
                        float up = __shfl_up_sync(0xffffffffu, cur, 1);
                        const float right = lds_f32(xs_row + o1);
                        const float down = lds_f32(xs_row + 512u + o);
                        const float x0v = lds_f32(x0_row + o);
                        if (lane == 0) up = lds_f32(xs_top + o);
                        const float newv = gs_update(x0v, right, cur, down, up, a, c_recip);
                        if (in_row) sts_f32(xs_row + om1, cur);                // column c-1 is final: F = R
                        if (lane == 31) sts_f32(raw_s + o, newv);
                        prev2 = cur;
                        prev_up = up;
                        cur = newv;
                        om1 = o;
                        o = o1;
                        __syncwarp();
                    }
                } else {
                    int c = 32 * m - lane;
#pragma unroll 2
                    for (int t = 0; t < 32; ++t, ++c) {
                        const uint32_t o1 = (o + 4u) & 508u;
                        float up = __shfl_up_sync(0xffffffffu, cur, 1);
                        const float right = lds_f32(xs_row + o1);
                        const float down = lds_f32(xs_row + 512u + o);
                        const float x0v = lds_f32(x0_row + o);
                        if (lane == 0) up = lds_f32(xs_top + o);
                        float newv = gs_update(x0v, right, cur, down, up, a, c_recip);
                        float F = cur;                                         // R_k(c-1, j)
                        if (ORIENT == EQ_ADJUST_ROW) {
                            const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 3u;
                            F = (code == EQ_CODE_ROW_RIGHT) ? -newv : ((code == EQ_CODE_ROW_LEFT) ? -prev2 : cur);
                        } else if (ORIENT == EQ_ADJUST_COLUMN) {
                            if (!in_row) newv = lds_f32(xs_row + o);           // frame row N-1 passes through
                            const float dn = __shfl_down_sync(0xffffffffu, newv, 1);
                            const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 12u;
                            if (code == EQ_CODE_COL_UP) F = -prev_up;
                            else if (code == EQ_CODE_COL_DOWN) {
                                if (lane < 31) F = -dn;
                                else if (last_band) F = -lds_f32(xs_top + 33u * 512u + om1);
                                // else: the band below patches this cell
                            }
                            if (lane == 0 && b > 0) {
                                // cell (c, j0-1) of the band above takes -R_k(c, j0) when its code says DOWN
                                const unsigned code0 = lds_u8(cs_top + (o >> 2)) & 12u;
                                if (code0 == EQ_CODE_COL_DOWN) x[(size_t)(j0 - 1) * P + c] = -newv;
                            }
                        } else {
                            // Passive band that owns a frame row (fluid.rs:182-183, conditional per Q6)
                            if ((j == 1 || j == N - 2) && lds_u8(cs_top + (om1 >> 2))) {
                                if (j == 1) sts_f32(xs_top + om1, cur);
                                if (j == N - 2) sts_f32(xs_row + 512u + om1, cur);
                            }
                        }
                        if (in_row) sts_f32(xs_row + om1, F);
                        if (lane == 31 && in_row) sts_f32(raw_s + o, newv);
                        prev2 = cur;
                        prev_up = up;
                        cur = newv;
                        om1 = o;
                        o = o1;
                        __syncwarp();
                    }
                }
            } else {
                // ---- general loop: frame columns, fix-ups, Passive frame copies ----------------------
                const int s_end = min(32 * m + 32, S);
                for (int s = 32 * m; s < s_end; ++s) {
                    const int c = s - lane;          // column this lane computes now (0 = left frame cell)
                    const uint32_t o = ((uint32_t)c & 127u) << 2;
                    const uint32_t om1 = (o - 4u) & 508u;
                    const float up = __shfl_up_sync(0xffffffffu, cur, 1);
                    float newv = cur;
                    float top = up;
                    if (c >= 0 && c <= N - 1) {
                        if (in_row && c >= 1 && c <= N - 2) {
                            const float right = lds_f32(xs_row + ((o + 4u) & 508u));
                            const float down = lds_f32(xs_row + 512u + o);
                            if (lane == 0) top = lds_f32(xs_top + o);
                            const float x0v = lds_f32(x0_row + o);
                            newv = gs_update(x0v, right, cur, down, top, a, c_recip);
                        } else if (in_row || ORIENT == EQ_ADJUST_COLUMN) {
                            newv = lds_f32(xs_row + o);      // frame column / frame row N-1: pass through
                        }
                    }
                    float dn = 0.f;
                    if (ORIENT == EQ_ADJUST_COLUMN) dn = __shfl_down_sync(0xffffffffu, newv, 1);
                    const int cf = c - 1;            // column finalised now
                    if (in_row && cf >= 1 && cf <= N - 2) {
                        float F = cur;               // R_k(cf, j)
                        if (ORIENT == EQ_ADJUST_ROW) {
                            const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 3u;
                            if (code == EQ_CODE_ROW_RIGHT) F = -newv;
                            else if (code == EQ_CODE_ROW_LEFT) F = -prev2;
                        } else if (ORIENT == EQ_ADJUST_COLUMN) {
                            const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 12u;
                            if (code == EQ_CODE_COL_UP) F = -prev_up;
                            else if (code == EQ_CODE_COL_DOWN) {
                                if (lane < 31) F = -dn;
                                else if (last_band) F = -lds_f32(xs_top + 33u * 512u + om1);
                                // else: the band below patches this cell (see below)
                            }
                        }
                        sts_f32(xs_row + om1, F);
                        if (ORIENT == EQ_PASSIVE) {   // fluid.rs:179-187, conditional per quirk Q6
                            if (row_has_fluid) {
                                if (cf == 1) sts_f32(xs_row, cur);
                                if (cf == N - 2) sts_f32(xs_row + (((uint32_t)(N - 1) & 127u) << 2), cur);
                            }
                            if ((j == 1 || j == N - 2) && lds_u8(cs_top + (om1 >> 2))) {
                                if (j == 1) sts_f32(xs_top + om1, cur);
                                if (j == N - 2) sts_f32(xs_row + 512u + om1, cur);
                            }
                        }
                    }
                    if (in_row && c >= 1 && c <= N - 2) {
                        if (ORIENT == EQ_ADJUST_COLUMN && lane == 0 && b > 0) {
                            // cell (c, j0-1) of the band above takes -R_k(c, j0) when its code says DOWN
                            const unsigned code0 = lds_u8(cs_top + (o >> 2)) & 12u;
                            if (code0 == EQ_CODE_COL_DOWN) x[(size_t)(j0 - 1) * P + c] = -newv;
                        }
                        if (lane == 31) sts_f32(raw_s + o, newv);
                    }
                    prev2 = cur;
                    prev_up = top;
                    cur = newv;
                    __syncwarp();
                }
            }
            // columns < 32m are final for every lane: hand chunk m-1 to the storer
            if (m >= 1 && m - 1 < NC && lane == 0) mbar_arrive(bar_done(m - 1));
        }
        if (lane == 0)
            for (int q = max(M - 1, 0); q < NC; ++q) mbar_arrive(bar_done(q));
        return true;
    }
};

// Three warps per CTA (compute / loader / storer); persistent CTAs pull jobs from the ticket counter.
__global__ void __launch_bounds__(LSX_THREADS) k_linsolve_exact(const LsxParams p) {
    EQ_DYN_SMEM(lsx_smem_raw);
    const uint32_t sbase = smem_u32(lsx_smem_raw);
    const int total = p.njobs * p.nprob;
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    for (;;) {
        __syncthreads();                               // the previous job is finished in all three roles
        if (threadIdx.x == 0) {
            const unsigned t = (ld_volatile_s32(p.error) != 0) ? 0xffffffffu : atomicAdd(p.ticket, 1u);
            sts_u32(sbase + LSX_MISC_OFF, t);
            for (int i = 0; i < 4; ++i) {
                mbar_init(sbase + LSX_BAR_OFF + (uint32_t)i * 16u, 32u);          // full: 32 loader lanes
                mbar_init(sbase + LSX_BAR_OFF + 64u + (uint32_t)i * 16u, 1u);     // done: compute lane 0
                mbar_init(sbase + LSX_BAR_OFF + 128u + (uint32_t)i * 16u, 1u);    // free: storer lane 0
            }
        }
        __syncthreads();
        const unsigned t = lds_u32(sbase + LSX_MISC_OFF);
        if (t >= (unsigned)total) break;
        const int pi = (int)(t % (unsigned)p.nprob);
        const uint32_t jb = p.jobs[t / (unsigned)p.nprob];
        const int k = (int)(jb >> 16), b = (int)(jb & 0xffffu);
        const LsxProblem &pr = p.prob[pi];
#define LSX_DISPATCH(O)                                           \
    {                                                             \
        const LsxJob<O> job(p, pr, sbase, b, k, lane);            \
        if (warp == 0) job.run_compute();                         \
        else if (warp == 1) job.run_loader();                     \
        else job.run_storer();                                    \
    }
        if (pr.orient == EQ_ADJUST_ROW) LSX_DISPATCH(EQ_ADJUST_ROW)
        else if (pr.orient == EQ_ADJUST_COLUMN) LSX_DISPATCH(EQ_ADJUST_COLUMN)
        else LSX_DISPATCH(EQ_PASSIVE)
#undef LSX_DISPATCH
        // a role that gave up has set *p.error (or seen it set); the next pass of the loop makes
        // thread 0 read it and every thread leaves together
    }
}
