// k_linsolve_tb.cuh -- the bit-exact wavefront solver of k_linsolve_exact.cuh with TEMPORAL
// BLOCKING: one job performs TBX_T consecutive Gauss-Seidel iterations of its band from the
// shared-memory tile before anything goes back to global memory.
//
// Job (b, g) = band b, iterations k0 = g*T .. k0+nsub-1.  The compute warp runs `nsub` sub-steps
// per step: sub-step t is iteration k0+t on the rows  j0-2t .. j0-2t+31  (the band moves up two
// rows per iteration) at column  s - lane - LAG*t.  Moving up by two rows is what makes every
// input of sub-step t+1 come from sub-step t of the SAME job or from the band ABOVE:
//   right  F_{k0+t}(c+1, j)   row j   = lane r-2 of sub-step t   (finalised LAG steps earlier)
//   down   F_{k0+t}(c,   j+1) row j+1 = lane r-1 of sub-step t
//   rows j0-2t-2, j0-2t-1 (lanes 0,1 of sub-step t+1) = lanes 30,31 of sub-step t of band b-1
//          ("edge rows", passed through global memory like the raw stream)
//   up     R_{k0+t+1}(c, j-1) by __shfl_up, lane 0 from band b-1's raw stream of sub-step t+1
// so the job still depends only on (b-1, g) and (b+1, g-1), but there are T times fewer jobs,
// global hand-offs and HBM/L2 bytes per iteration, and the compute warp interleaves T
// independent dependency chains (the SHFL -> 4 FP latency of one hides behind the others).
// The tile is updated in place exactly like the reference's array: sub-step t+1 overwrites
// F_{k0+t} with F_{k0+t+1} LAG columns behind sub-step t.
//
// Global x after a group holds iteration k0+nsub-1 for every row: band b stores the rows of
// its last sub-step; NB' = ceil((N-2+SK)/32) bands cover the shifted row axis (SK = 2(T-1)).
// Single GPU only (the row-slab path keeps using k_linsolve_exact).
#pragma once
#include "k_linsolve_exact.cuh"
#include <type_traits>

#ifndef TBX_T
#define TBX_T 2                                   // iterations per job
#endif
// TBX_SPLIT = 1: one compute warp PER SUB-STEP.  The per-job limit of the fused kernel is the instruction stream of
// its single compute warp (~62 instructions per step for two sub-steps; one warp issues well below one instruction per
// cycle), not the latency of the SHFL -> FP chain.  With the sub-steps a whole chunk apart (TBX_LAG >= LSX_CW + 2)
// everything sub-step t+1 reads in macro step m was finalised by sub-step t in macro step m-1 at the latest, and what
// either writes in a macro step lies outside the columns the other reads in it, so the two warps run the same macro
// step concurrently and only meet at a named barrier between macro steps.
#ifndef TBX_SPLIT
#define TBX_SPLIT 0
#endif
#ifndef TBX_LAG
#define TBX_LAG (TBX_SPLIT ? (EQ_LSX_CW + 2) : 2)  // columns between consecutive sub-steps
#endif
static_assert(!TBX_SPLIT || (TBX_T == 2 && TBX_LAG >= EQ_LSX_CW + 2), "TBX_SPLIT: two sub-steps, a whole chunk apart");
#define TBX_CWARPS (TBX_SPLIT ? TBX_T : 1)         // compute warps per CTA
#define TBX_THREADS (32 * (3 + TBX_CWARPS))        // + loader, storer, publisher
#ifndef TBX_SLOTS
#define TBX_SLOTS 8                               // chunks in the staging ring (power of two)
#endif
#define TBX_RING (TBX_SLOTS * LSX_CW + 0u)        // columns in the ring = bytes per code row
#define TBX_ROWB (TBX_RING * 4u)                  // bytes per tile row
#define TBX_OMASK (TBX_ROWB - 4u)                 // byte-offset mask inside a row
#define TBX_CMASK (TBX_RING - 1u)                 // column mask
#define TBX_SK (2 * (TBX_T - 1))                  // rows the band has moved up at its last sub-step
#define TBX_XROWS (34 + TBX_SK)                   // tile rows: global rows j0-SK-1 .. j0+32
#define TBX_X0ROWS (32 + TBX_SK)                  // global rows j0-SK .. j0+31
#define TBX_BACK ((32 + TBX_LAG * (TBX_T - 1) + 1 + LSX_CW - 1) / LSX_CW)   // after macro m, chunk m-BACK is final
#define TBX_XS_OFF 0u
#define TBX_X0_OFF (TBX_XROWS * TBX_ROWB)
#define TBX_CS_OFF (TBX_X0_OFF + TBX_X0ROWS * TBX_ROWB)
#define TBX_RAWIN_OFF (TBX_CS_OFF + TBX_XROWS * TBX_RING)     // T rings: R of the row above, per sub-step
#define TBX_RAWOUT_OFF (TBX_RAWIN_OFF + TBX_T * TBX_ROWB)     // T rings: R of my last row, per sub-step
#define TBX_BAR_OFF (TBX_RAWOUT_OFF + TBX_T * TBX_ROWB)
#define TBX_MISC_OFF (TBX_BAR_OFF + 3u * TBX_SLOTS * 16u)
#define TBX_AB_OFF (TBX_MISC_OFF + 16u)                       // mbarrier the compute warps meet at between macro steps
#define TBX_SMEM_BYTES (TBX_AB_OFF + 16u)

struct TbxProblem {
    float *x;
    const float *x0;
    float *raw;          // [T][NBP][P]    R of the last row of band b-1 at sub-step t, read by band b
    float *edge;         // [T-1][NBP][2][P] F of the last two rows of band b-1 at sub-step t, read by band b at t+1
    unsigned *progress;  // [G][NBP] chunks completed
    float a, c_recip;
    int orient;
};

struct TbxParams {
    TbxProblem prob[2];
    int nprob;
    const uint8_t *codes;
    const uint8_t *chunk_flags;  // [2][NBP][NC]: rows j0-SK-1 .. j0+31 of band b hold [0] a Row code, [1] a Column code other
                                 // than the UP of row 1 / DOWN of row N-2 every column has (k_build_codes)
    const uint8_t *row_fluid;
    const uint8_t *col_fluid;
    const uint32_t *jobs;        // [G*NBP] (g << 16 | b) in wavefront order
    int njobs;
    int N, P, K, G, NBP, NC;     // K iterations in G groups of TBX_T (the last one may be shorter)
    unsigned *ticket;
    int *error;
    int pub_batch;               // the publisher releases progress every pub_batch chunks (and at the end): one release
                                 // costs ~4 us of fence whatever it covers; one chunk per release paces the whole chain
    int rotate_roles;            // see LsxParams
    int debug_nodeps;            // EQ_LSX_NODEPS=1: skip the dependency waits (WRONG results; throughput experiments only)
    const unsigned *run_if;      // not null: return at once when *run_if == 0 (see LsxParams::run_if)
    int single_shot;             // one job per CTA (see LsxParams::single_shot)
    int passive_fast_frames;     // every interior column has a NoWall cell: Passive frame-row copies are unconditional
    int trace_g, trace_b, trace_q;  // EQ_LSX_TRACE=g,b,q: group, first band and first chunk traced
    unsigned long long *trace;    // optional event trace (EQ_LSX_TRACE=1): [4 bands][8 events][128 chunks] ns, see dump_lsx_stats
    unsigned long long *jobtimes; // optional [4 * njobs] (problem 0): start, first chunk ready, end (ns), SM id (EQ_LSX_JOBTIMES=file)
};

#define TBX_TRACE(ev, q) do { if (p.trace && lane == 0 && g == p.trace_g && b >= p.trace_b && b < p.trace_b + 4 && (q) >= p.trace_q && (q) < p.trace_q + 128) p.trace[((size_t)(b - p.trace_b) * 8 + (ev)) * 128 + (q) - p.trace_q] = lsx_gtime(); } while (0)

template <int ORIENT>
struct TbxJob {
    const TbxParams &p;
    const TbxProblem &pr;
    uint32_t sbase;
    int b, g, lane;
    int N, P, NC, NBP, j0, M, nsub, k0;
    const uint8_t *cflags;

    __device__ __forceinline__ TbxJob(const TbxParams &p_, const TbxProblem &pr_, uint32_t sbase_, int b_, int g_, int lane_)
        : p(p_), pr(pr_), sbase(sbase_), b(b_), g(g_), lane(lane_) {
        N = p.N; P = p.P; NC = p.NC; NBP = p.NBP;
        j0 = 1 + 32 * b;
        k0 = g * TBX_T;
        nsub = min(TBX_T, p.K - k0);
        M = (N + 31 + TBX_LAG * (TBX_T - 1) + LSX_CW - 1) / LSX_CW;
        cflags = p.chunk_flags + (ORIENT == EQ_ADJUST_COLUMN ? (size_t)NBP * NC : 0) + (size_t)b * NC;
    }
    __device__ __forceinline__ uint32_t bar_full(int q) const { return sbase + TBX_BAR_OFF + (uint32_t)(q % TBX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t bar_done(int q) const { return sbase + TBX_BAR_OFF + (uint32_t)(TBX_SLOTS + q % TBX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t bar_free(int q) const { return sbase + TBX_BAR_OFF + (uint32_t)(2 * TBX_SLOTS + q % TBX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t use_parity(int q) const { return (uint32_t)((q / TBX_SLOTS) & 1); }
    // tile row of a global row
    __device__ __forceinline__ int trow(int j) const { return j - (j0 - TBX_SK - 1); }
    // does any sub-step of this band touch the first or the last interior row (Passive frame-row copies)?
    __device__ __forceinline__ bool owns_frame_row() const { return b == 0 || (j0 + 31 >= N - 2 && j0 - TBX_SK <= N - 2); }

    enum { MODE_FAST = 0, MODE_CODED = 1, MODE_EDGE = 2 };
    __device__ __forceinline__ int macro_mode(int m) const {
        if (LSX_CW * m - 31 - TBX_LAG * (TBX_T - 1) < 2 || LSX_CW * m + LSX_CW - 1 > N - 2) return MODE_EDGE;
        // Passive has no per-cell codes; the bands next to the frame rows copy their first / last row into
        // the frame (quirk Q6: only where the column holds a NoWall cell -- when every column does, the
        // fast loop does the copy unconditionally, otherwise the general loop looks col_fluid up).
        // These bands head the dependency chain of their group: they must not be slower than the rest.
        if (ORIENT == EQ_PASSIVE) return (owns_frame_row() && !p.passive_fast_frames) ? MODE_CODED : MODE_FAST;
        unsigned any = 0;
#pragma unroll
        for (int d = 0; d <= TBX_BACK; ++d) any |= cflags[m - d];
        return any ? MODE_CODED : MODE_FAST;
    }
    // AdjustColumn: may an EDGE macro step use the fast loop (no Column code but the frame-adjacent ones)?
    __device__ __forceinline__ bool edge_simple(int m) const {
        if (ORIENT != EQ_ADJUST_COLUMN) return true;
        unsigned any = 0;
#pragma unroll
        for (int d = 0; d <= TBX_BACK; ++d) any |= cflags[min(max(m - d, 0), NC - 1)];
        return any == 0;
    }
    __device__ __forceinline__ bool need_codes(int q) const {
        if (ORIENT == EQ_PASSIVE) {                                  // row 0 of the code tile carries col_fluid
            if (!owns_frame_row()) return false;
            if (!p.passive_fast_frames) return true;
        }
        bool need = false;
#pragma unroll
        for (int d = 0; d <= TBX_BACK; ++d) need = need || (macro_mode(q + d) != MODE_FAST);
        return need;
    }

    // ------------------------------------------------------------------ LOADER warp
    __device__ __forceinline__ bool run_loader() const {
        const float *__restrict__ x = pr.x;
        const float *__restrict__ x0 = pr.x0;
        const unsigned *flag_prev = (g > 0 && !p.debug_nodeps) ? pr.progress + (size_t)(g - 1) * NBP + min(b + 1, NBP - 1) : nullptr;
        const unsigned *flag_above = (b > 0 && !p.debug_nodeps) ? pr.progress + (size_t)g * NBP + (b - 1) : nullptr;
        constexpr int LPR = LSX_CW / 4, RPP = 32 / LPR;
        const int sub = lane % LPR, rr = lane / LPR;
        for (int q = 0; q < NC; ++q) {
            if (q + LSX_PF < NC) prefetch_l2(x0 + (size_t)(j0 + lane) * P + LSX_CW * (q + LSX_PF));
            if (q >= TBX_SLOTS && !lsx_wait_bar(bar_free(q), use_parity(q - TBX_SLOTS), p.error, lane)) return false;
            TBX_TRACE(0, q);
            if (!lsx_wait_flags(flag_prev, (unsigned)q + 1u, false, flag_above, (unsigned)q + 1u, false, p.error, lane)) return false;
            const uint32_t slot = (uint32_t)(q % TBX_SLOTS) * (LSX_CW * 4u);
            const int col0 = LSX_CW * q;
#ifdef TBX_DBG_NOLOAD   // stage-isolation build (with EQ_LSX_NODEPS=1; WRONG results): the loader only signals
            (void)x; (void)x0; (void)slot; (void)col0; (void)sub; (void)rr;
            if (false)
#endif
            {
            // x: global rows j0 .. j0+32 hold F_{k0-1} (band 0 also needs the frame row 0)
            const int jx_lo = (b == 0) ? 0 : j0;
#pragma unroll
            for (int pass = 0; pass < (34 + RPP - 1) / RPP; ++pass) {
                const int j = jx_lo + RPP * pass + rr;
                if (j <= j0 + 32)
                    cp_async_16s(sbase + TBX_XS_OFF + (uint32_t)trow(j) * TBX_ROWB + slot + 16u * sub, x + (size_t)j * P + col0 + 4 * sub);
            }
            if (b > 0) {
                // edge rows: F_{k0+t-1} of global rows j0-2t, j0-2t+1 for sub-step t >= 1; raw tops for every sub-step
                const int task = lane / LPR;                     // RPP tasks per pass
#pragma unroll
                for (int pass = 0; pass < (2 * (TBX_T - 1) + TBX_T + RPP - 1) / RPP; ++pass) {
                    const int id = RPP * pass + task;
                    if (id < 2 * (nsub - 1)) {
                        const int t = 1 + id / 2, which = id & 1;
                        const float *src = pr.edge + (((size_t)(t - 1) * NBP + b) * 2 + which) * P + col0 + 4 * sub;
                        cp_async_16s(sbase + TBX_XS_OFF + (uint32_t)trow(j0 - 2 * t + which) * TBX_ROWB + slot + 16u * sub, src);
                    } else if (id >= 2 * (TBX_T - 1) && id - 2 * (TBX_T - 1) < nsub) {
                        const int t = id - 2 * (TBX_T - 1);
                        cp_async_16s(sbase + TBX_RAWIN_OFF + (uint32_t)t * TBX_ROWB + slot + 16u * sub,
                                     pr.raw + ((size_t)t * NBP + b) * P + col0 + 4 * sub);
                    }
                }
            }
#pragma unroll
            for (int pass = 0; pass < (TBX_X0ROWS + RPP - 1) / RPP; ++pass) {
                const int i = RPP * pass + rr;                   // x0 tile row i = global row j0-SK+i
                const int j = j0 - TBX_SK + i;
                if (i < TBX_X0ROWS && j >= 0)
                    cp_async_16s(sbase + TBX_X0_OFF + (uint32_t)i * TBX_ROWB + slot + 16u * sub, x0 + (size_t)j * P + col0 + 4 * sub);
            }
            if (need_codes(q)) {
                constexpr int CLPR = LSX_CW / 16;
                const uint32_t cslot = (uint32_t)(q % TBX_SLOTS) * LSX_CW;
                if (ORIENT == EQ_PASSIVE) {
                    if (lane < CLPR) cp_async_16s(sbase + TBX_CS_OFF + cslot + 16u * lane, p.col_fluid + col0 + 16 * lane);
                } else {
                    const int csub = lane % CLPR, crr = lane / CLPR;
#pragma unroll
                    for (int pass = 0; pass < (TBX_XROWS * CLPR + 31) / 32; ++pass) {
                        const int t = (32 / CLPR) * pass + crr;  // code tile row t = global row j0-SK-1+t
                        const int j = j0 - TBX_SK - 1 + t;
                        if (t < TBX_XROWS - 1 && j >= 0)
                            cp_async_16s(sbase + TBX_CS_OFF + (uint32_t)t * TBX_RING + cslot + 16u * csub,
                                         p.codes + (size_t)j * P + col0 + 16 * csub);
                    }
                }
            }
            }
            cp_async_mbar_arrive_noinc(bar_full(q));
            TBX_TRACE(1, q);
        }
        return true;
    }

    // ------------------------------------------------------------------ STORER warp
    __device__ __forceinline__ bool run_storer() const {
        float *__restrict__ x = pr.x;
        constexpr int LPR = LSX_CW / 4, RPP = 32 / LPR;
        const int sub = lane % LPR, rr = lane / LPR;
        const int jf = j0 - 2 * (nsub - 1);                      // first row of the last sub-step
        const bool has_below = (b + 1 < NBP);
        for (int q = 0; q < NC; ++q) {
            if (!lsx_wait_bar(bar_done(q), use_parity(q), p.error, lane)) return false;
            TBX_TRACE(4, q);
            const uint32_t slot = (uint32_t)(q % TBX_SLOTS) * (LSX_CW * 4u);
            const int col0 = LSX_CW * q;
#ifdef TBX_DBG_NOSTORE  // stage-isolation build (WRONG results): the storer only recycles the slots
            (void)x; (void)slot; (void)col0; (void)sub; (void)rr; (void)jf; (void)has_below;
            if (false)
#endif
            {
            // rows of the last sub-step (iteration k0+nsub-1); Passive also keeps the frame rows it touched
#pragma unroll
            for (int pass = 0; pass < (34 + RPP - 1) / RPP; ++pass) {
                const int j = jf - 1 + RPP * pass + rr;          // jf-1 .. jf+32
                const bool interior = (j >= max(jf, 1) && j <= min(jf + 31, N - 2));
                const bool frame = (ORIENT == EQ_PASSIVE) && ((j == 0 && jf <= 1) || (j == N - 1 && jf <= N - 2 && jf + 31 >= N - 2));
                if (interior || frame) {
                    const float4 v = lds_f32x4(sbase + TBX_XS_OFF + (uint32_t)trow(j) * TBX_ROWB + slot + 16u * sub);
                    *reinterpret_cast<float4 *>(x + (size_t)j * P + col0 + 4 * sub) = v;
                }
            }
            if (has_below) {
                // edge rows (F of my rows 30,31 at sub-step t < nsub-1) and raw rows (R of my row 31) for band b+1
                const int task = lane / LPR;
#pragma unroll
                for (int pass = 0; pass < (2 * (TBX_T - 1) + TBX_T + RPP - 1) / RPP; ++pass) {
                    const int id = RPP * pass + task;
                    if (id < 2 * (nsub - 1)) {
                        const int t = id / 2, which = id & 1;
                        const float4 v = lds_f32x4(sbase + TBX_XS_OFF + (uint32_t)trow(j0 - 2 * t + 30 + which) * TBX_ROWB + slot + 16u * sub);
                        *reinterpret_cast<float4 *>(pr.edge + (((size_t)t * NBP + b + 1) * 2 + which) * P + col0 + 4 * sub) = v;
                    } else if (id >= 2 * (TBX_T - 1) && id - 2 * (TBX_T - 1) < nsub) {
                        const int t = id - 2 * (TBX_T - 1);
                        const float4 v = lds_f32x4(sbase + TBX_RAWOUT_OFF + (uint32_t)t * TBX_ROWB + slot + 16u * sub);
                        *reinterpret_cast<float4 *>(pr.raw + ((size_t)t * NBP + b + 1) * P + col0 + 4 * sub) = v;
                    }
                }
            }
            }
            __syncwarp();
            TBX_TRACE(5, q);
            if (lane == 0) {
                mbar_arrive(bar_free(q));
                sts_release_cta_u32(sbase + TBX_MISC_OFF + 4u, (uint32_t)q + 1u);
            }
        }
        return true;
    }

    // ------------------------------------------------------------------ PUBLISHER warp
    __device__ __forceinline__ bool run_publisher() const {
        unsigned *my_flag = pr.progress + (size_t)g * NBP + b;
        const uint32_t cnt = sbase + TBX_MISC_OFF + 4u;
        int q = 0, ok = 1;
        while (q < NC && ok) {
            if (lane == 0) {
                unsigned spins = 0;
                unsigned long long t0 = 0;
                int have;
                while ((have = (int)lds_acquire_cta_u32(cnt)) < min(q + p.pub_batch, NC)) {
                    __nanosleep(64);
                    if ((++spins & 1023u) == 0) {
                        if (lsx_expired(t0, spins)) { *p.error = 3; ok = 0; break; }
                        if (ld_volatile_s32(p.error) != 0) { ok = 0; break; }
                    }
                }
                if (ok) {
                    q = have;
                    TBX_TRACE(6, q - 1);
                    st_release_u32(my_flag, (unsigned)q);
                    TBX_TRACE(7, q - 1);
                }
            }
            q = __shfl_sync(0xffffffffu, q, 0);
            ok = __shfl_sync(0xffffffffu, ok, 0);
        }
        return ok != 0;
    }

    // ------------------------------------------------------------------ COMPUTE warp
    // TSEL < 0: this warp runs every sub-step; TSEL = t: only sub-step t (TBX_SPLIT)
    template <int TSEL>
    __device__ __forceinline__ bool run_compute() const {
        const float a = pr.a, c_recip = pr.c_recip;
        float *__restrict__ x = pr.x;
        const uint32_t xs0 = sbase + TBX_XS_OFF, cs0 = sbase + TBX_CS_OFF;
        const int S = N + 31 + TBX_LAG * (TBX_T - 1);
        // per sub-step: my row, its shared addresses, where the value above the first row comes from
        int row[TBX_T];
        bool in_row[TBX_T], first[TBX_T], rowfl[TBX_T], frame_from_edge[TBX_T], frame_top[TBX_T], frame_bot[TBX_T];
        uint32_t xs_row[TBX_T], x0_row[TBX_T], cs_row[TBX_T], top_base[TBX_T], down_row[TBX_T];
        float cur[TBX_T], prev2[TBX_T], prev_up[TBX_T];
#pragma unroll
        for (int t = 0; t < TBX_T; ++t) {
            row[t] = j0 - 2 * t + lane;
            in_row[t] = (t < nsub) && row[t] >= 1 && row[t] <= N - 2;
            first[t] = (b == 0) ? (row[t] == 1) : (lane == 0);
            rowfl[t] = (ORIENT == EQ_PASSIVE && in_row[t]) ? (p.row_fluid[row[t]] != 0) : false;
            frame_from_edge[t] = (ORIENT == EQ_PASSIVE) && t > 0 && b > 0 && lane == 1 && in_row[t] && row[t] == N - 2;
            frame_top[t] = (ORIENT == EQ_PASSIVE) && in_row[t] && row[t] == 1;
            frame_bot[t] = (ORIENT == EQ_PASSIVE) && in_row[t] && row[t] == N - 2;
            const int tr = trow(row[t]);                          // may be out of the tile for inactive lanes of band 0
            const int trc = min(max(tr, 0), TBX_XROWS - 2);
            xs_row[t] = xs0 + (uint32_t)trc * TBX_ROWB;
            // fast loops: where the lower neighbour is read (see frame_from_edge in the general loop)
            down_row[t] = frame_from_edge[t] ? xs_row[t] : xs_row[t] + TBX_ROWB;
            x0_row[t] = sbase + TBX_X0_OFF + (uint32_t)min(max(tr - 1, 0), TBX_X0ROWS - 1) * TBX_ROWB;
            cs_row[t] = cs0 + (uint32_t)trc * TBX_RING;
            top_base[t] = (b == 0) ? xs0 + (uint32_t)trow(0) * TBX_ROWB : sbase + TBX_RAWIN_OFF + (uint32_t)t * TBX_ROWB;
            cur[t] = prev2[t] = prev_up[t] = 0.f;
        }
        const uint32_t raw_out = sbase + TBX_RAWOUT_OFF;

        int m_cur = 0;
        const uint32_t last_col = ((uint32_t)(N - 1) & TBX_CMASK) << 2;
        // ---- fast loop: straight-line code for the 16 steps x T sub-steps of a macro step in which nothing needs
        // a per-cell code lookup.  The operands of step s+1 are fetched before the arithmetic of step s (nobody
        // writes them later than step s-1), so the step-to-step chain of a sub-step is SHFL -> 3 FADD, FMUL, FADD,
        // FMUL, and the T chains interleave.
        //   ZONE 0: every lane of every sub-step is on an interior column (MODE_FAST).
        //   ZONE 1 / 2: the macro steps at the start / end of the rows (MODE_EDGE), where lanes are still left of
        //   column 1 or already right of column N-2: the same loop plus range predicates (compares against
        //   immediates once unrolled), the frame-column pass-through, AdjustRow's codes (columns 1 and N-2 always
        //   carry one) and Passive's frame-column copies.  These steps open and close every job: the band below
        //   cannot start before the first ones are done nor finish before the last, so their duration is the lag
        //   between consecutive bands and the solve is a chain of NB such lags -- they must not be slow.
        const bool lean_edges = (ORIENT != EQ_PASSIVE) || !owns_frame_row() || p.passive_fast_frames;
        bool col_top[TBX_T], col_bot[TBX_T];
#pragma unroll
        for (int t = 0; t < TBX_T; ++t) {
            col_top[t] = (ORIENT == EQ_ADJUST_COLUMN) && in_row[t] && row[t] == 1;
            col_bot[t] = (ORIENT == EQ_ADJUST_COLUMN) && in_row[t] && row[t] == N - 2;
        }
        auto fast_steps = [&](auto zone_c) {
            constexpr int ZONE = decltype(zone_c)::value;
            const int cb = LSX_CW * m_cur - lane;                 // column of sub-step 0 at the first step
            const uint32_t ob = ((uint32_t)cb & TBX_CMASK) << 2;
            float right[TBX_T], down[TBX_T], x0v[TBX_T], topv[TBX_T];
#pragma unroll
            for (int t = 0; t < TBX_T; ++t) {
                if (TSEL >= 0 && t != TSEL) continue;
                const uint32_t o = (ob - 4u * TBX_LAG * t) & TBX_OMASK;
                right[t] = lds_f32(xs_row[t] + ((o + 4u) & TBX_OMASK));
                down[t] = lds_f32(down_row[t] + o);
                x0v[t] = lds_f32(x0_row[t] + o);
                topv[t] = first[t] ? lds_f32(top_base[t] + o) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < LSX_CW; ++i) {
                // Phase A: the shuffles and the operand fetches of BOTH sub-steps, back to back.  Nothing a sub-step
                // stores in this step is read by the other one before the next step (its inputs were finalised at
                // least one step earlier, see the header), so the T dependency chains SHFL -> FADD, FMUL, FADD, FMUL
                // are independent inside a step and the scheduler interleaves them -- as long as no volatile
                // shared-memory access of sub-step t sits between the arithmetic of t and the shuffle of t+1.
                float up[TBX_T], right_n[TBX_T], down_n[TBX_T], x0_n[TBX_T], self[TBX_T], below[TBX_T];
                unsigned code[TBX_T];
#pragma unroll
                for (int t = 0; t < TBX_T; ++t)
                    if (TSEL < 0 || t == TSEL) up[t] = __shfl_up_sync(0xffffffffu, cur[t], 1);
#pragma unroll
                for (int t = 0; t < TBX_T; ++t) {
                    if (TSEL >= 0 && t != TSEL) continue;
                    const uint32_t o = (ob + 4u * i - 4u * TBX_LAG * t) & TBX_OMASK;
                    const uint32_t o1 = (o + 4u) & TBX_OMASK, o2 = (o + 8u) & TBX_OMASK, om1 = (o - 4u) & TBX_OMASK;
                    right_n[t] = lds_f32(xs_row[t] + o2);
                    down_n[t] = lds_f32(down_row[t] + o1);
                    x0_n[t] = lds_f32(x0_row[t] + o1);
                    if (first[t]) {
                        up[t] = topv[t];
                        topv[t] = lds_f32(top_base[t] + o1);
                    }
                    self[t] = 0.f;
                    code[t] = 0u;
                    below[t] = 0.f;
                    if (ZONE != 0) {
                        self[t] = lds_f32(xs_row[t] + o);
                        if (ORIENT == EQ_ADJUST_ROW) code[t] = lds_u8(cs_row[t] + (om1 >> 2)) & 3u;
                    }
                    // rows 1 and N-2 mirror the frame rows (which AdjustColumn never changes, quirk Q5):
                    // x[i,1] = -x[i,0] (the `up` this cell was computed with), x[i,N-2] = -x[i,N-1]
                    if (ORIENT == EQ_ADJUST_COLUMN) below[t] = lds_f32(xs_row[t] + TBX_ROWB + om1);
                }
                // Phase B: arithmetic and the stores of the step
#pragma unroll
                for (int t = 0; t < TBX_T; ++t) {
                    if (TSEL >= 0 && t != TSEL) continue;
                    const int c = cb + i - TBX_LAG * t;
                    const uint32_t o = (ob + 4u * i - 4u * TBX_LAG * t) & TBX_OMASK;
                    const uint32_t om1 = (o - 4u) & TBX_OMASK;
                    float newv = gs_update(x0v[t], right[t], cur[t], down[t], up[t], a, c_recip);
                    bool interior = in_row[t], fin = in_row[t];
                    float F = cur[t];
                    if (ZONE != 0) {
                        const bool frame_col = in_row[t] & (ZONE == 1 ? (c == 0) : (c == N - 1));
                        interior = in_row[t] & (ZONE == 1 ? (c >= 1) : (c <= N - 2));
                        fin = in_row[t] & (ZONE == 1 ? (c >= 2) : (c <= N - 1));        // cf = c-1 in 1 .. N-2
                        newv = interior ? newv : (frame_col ? self[t] : cur[t]);
                        if (ORIENT == EQ_ADJUST_ROW)
                            F = (code[t] == EQ_CODE_ROW_RIGHT) ? -newv : ((code[t] == EQ_CODE_ROW_LEFT) ? -prev2[t] : cur[t]);
                    }
                    if (ORIENT == EQ_ADJUST_COLUMN) F = col_top[t] ? -prev_up[t] : (col_bot[t] ? -below[t] : F);
                    if (fin) sts_f32(xs_row[t] + om1, F);
                    if (ORIENT == EQ_PASSIVE) {
                        // fluid.rs:182-186 (every col_fluid is set here: passive_fast_frames)
                        if (ZONE == 1) { if (fin & rowfl[t] & (c == 2)) sts_f32(xs_row[t], cur[t]); }
                        if (ZONE == 2) { if (fin & rowfl[t] & (c == N - 1)) sts_f32(xs_row[t] + last_col, cur[t]); }
                        if (fin & frame_top[t]) sts_f32(xs_row[t] - TBX_ROWB + om1, cur[t]);
                        if (fin & frame_bot[t]) sts_f32(xs_row[t] + TBX_ROWB + om1, cur[t]);
                    }
                    if (interior & (lane == 31)) sts_f32(raw_out + (uint32_t)t * TBX_ROWB + o, newv);
                    prev2[t] = cur[t];
                    prev_up[t] = up[t];
                    cur[t] = newv;
                    right[t] = right_n[t];
                    down[t] = down_n[t];
                    x0v[t] = x0_n[t];
                }
                __syncwarp();
            }
        };

        // ---- general loop: macro steps that touch fix-up codes (MODE_CODED) or run off the ends of the rows
        // (MODE_EDGE, RANGED).  Flat on purpose -- selects and single-level predicated stores: every shared
        // address is valid for any column (the ring wraps), so operands are loaded unconditionally and the
        // tests only pick results.  (A version with nested ifs compiled to divergent branches and took ~10 us
        // per macro step against 1.1 us for the fast loop.  The EDGE steps open and close every job: the band
        // below cannot start before the first ones are done nor finish before the last, so their duration is
        // the lag between consecutive bands and the solve is a chain of NB such lags.)
        auto general_steps = [&](auto ranged_c) {
            constexpr bool RANGED = decltype(ranged_c)::value;
            const int s_end = min(LSX_CW * m_cur + LSX_CW, S);
            for (int s = LSX_CW * m_cur; s < s_end; ++s) {
#pragma unroll
                for (int t = 0; t < TBX_T; ++t) {
                    if (TSEL >= 0 && t != TSEL) continue;
                    if (t < nsub) {
                        const int c = s - lane - TBX_LAG * t;       // column this lane computes now (0 = left frame cell)
                        const int cf = c - 1;                        // column finalised now
                        const uint32_t o = ((uint32_t)c & TBX_CMASK) << 2;
                        const uint32_t om1 = (o - 4u) & TBX_OMASK;
                        const float up = __shfl_up_sync(0xffffffffu, cur[t], 1);
                        const float right = lds_f32(xs_row[t] + ((o + 4u) & TBX_OMASK));
                        float down = lds_f32(xs_row[t] + TBX_ROWB + o);
                        const float x0v = lds_f32(x0_row[t] + o);
                        const float self = lds_f32(xs_row[t] + o);
                        const float top = first[t] ? lds_f32(top_base[t] + o) : up;
                        if (ORIENT == EQ_PASSIVE) {
                            // Row N-2 arrived as band b-1's LAST row of the previous sub-step: that band copied it into
                            // the frame row N-1 of ITS tile (where col_fluid), mine still holds the older frame.
                            const bool colf_c = lds_u8(cs0 + (o >> 2)) != 0;
                            down = (frame_from_edge[t] & colf_c) ? self : down;
                        }
                        const bool gs_ok = in_row[t] & (!RANGED | ((c >= 1) & (c <= N - 2)));
                        const bool pass_row = in_row[t] | ((ORIENT == EQ_ADJUST_COLUMN) & (row[t] == N - 1));
                        const bool pass_ok = pass_row & (!RANGED | ((c >= 0) & (c <= N - 1)));
                        const float gval = gs_update(x0v, right, cur[t], down, top, a, c_recip);
                        const float newv = gs_ok ? gval : (pass_ok ? self : cur[t]);
                        const bool fin = in_row[t] & (!RANGED | ((cf >= 1) & (cf <= N - 2)));
                        float F = cur[t];
                        if (ORIENT == EQ_ADJUST_ROW) {
                            const unsigned code = lds_u8(cs_row[t] + (om1 >> 2)) & 3u;
                            F = (code == EQ_CODE_ROW_RIGHT) ? -newv : ((code == EQ_CODE_ROW_LEFT) ? -prev2[t] : cur[t]);
                        } else if (ORIENT == EQ_ADJUST_COLUMN) {
                            const float dn = __shfl_down_sync(0xffffffffu, newv, 1);
                            const unsigned code = lds_u8(cs_row[t] + (om1 >> 2)) & 12u;
                            // lane 31's lower neighbour is either the frame row N-1 (in the tile) or the first
                            // row of band b+1 at this sub-step, which patches the cell itself (below)
                            const bool below_is_frame = (row[t] == N - 2);
                            const float below = (lane < 31) ? dn : lds_f32(xs_row[t] + TBX_ROWB + om1);
                            const bool take_down = (code == EQ_CODE_COL_DOWN) & ((lane < 31) | below_is_frame);
                            F = (code == EQ_CODE_COL_UP) ? -prev_up[t] : (take_down ? -below : cur[t]);
                            // cell (c, j-1) above my first row belongs to band b-1 at this sub-step: it takes -R(c, j)
                            // when its code says DOWN.  For an intermediate iteration the cell lives on in MY tile
                            // (it is one of my edge rows); for the last one it is already in global x.
                            const unsigned code0 = lds_u8(cs_row[t] - TBX_RING + (o >> 2)) & 12u;
                            const bool patch = gs_ok & (lane == 0) & (b > 0) & (code0 == EQ_CODE_COL_DOWN);
                            if (patch & (t == nsub - 1)) x[(size_t)(row[t] - 1) * P + c] = -newv;
                            if (patch & (t != nsub - 1)) sts_f32(xs_row[t] - TBX_ROWB + o, -newv);
                        }
                        if (fin) sts_f32(xs_row[t] + om1, F);
                        if (ORIENT == EQ_PASSIVE) {
                            // fluid.rs:182-186, conditional per quirk Q6 (col_fluid is staged in row 0 of the code tile)
                            if (RANGED) {
                                if (fin & rowfl[t] & (cf == 1)) sts_f32(xs_row[t], cur[t]);
                                if (fin & rowfl[t] & (cf == N - 2)) sts_f32(xs_row[t] + last_col, cur[t]);
                            }
                            const bool colf = lds_u8(cs0 + (om1 >> 2)) != 0;
                            if (fin & colf & (row[t] == 1)) sts_f32(xs_row[t] - TBX_ROWB + om1, cur[t]);
                            if (fin & colf & (row[t] == N - 2)) sts_f32(xs_row[t] + TBX_ROWB + om1, cur[t]);
                        }
                        if (gs_ok & (lane == 31)) sts_f32(raw_out + (uint32_t)t * TBX_ROWB + o, newv);
                        prev2[t] = cur[t];
                        prev_up[t] = top;
                        cur[t] = newv;
                        __syncwarp();
                    }
                }
            }
        };

        if (!lsx_wait_bar(bar_full(0), 0u, p.error, lane)) return false;
        if (TSEL <= 0 && p.jobtimes && lane == 0 && &pr == &p.prob[0]) p.jobtimes[4 * ((size_t)g * NBP + b) + 1] = lsx_gtime();
        for (int m = 0; m < M; ++m) {
            if (m + 1 < NC && !lsx_wait_bar(bar_full(m + 1), use_parity(m + 1), p.error, lane)) return false;
            TBX_TRACE(2, m);
            const int mode = macro_mode(m);
            m_cur = m;
            const int s_end = min(LSX_CW * m + LSX_CW, S);
#ifdef TBX_DBG_NOCOMPUTE   // stage-isolation build (WRONG results): the compute warp only waits and signals
            if (false) {
            } else if (mode >= 0 && s_end >= 0) {
            } else
#endif
            if (nsub == TBX_T && mode == MODE_FAST) {
                fast_steps(std::integral_constant<int, 0>{});
            } else if (nsub == TBX_T && mode == MODE_EDGE && lean_edges && edge_simple(m) && LSX_CW * m + LSX_CW - 1 <= N - 2) {
                fast_steps(std::integral_constant<int, 1>{});       // start of the rows only
            } else if (nsub == TBX_T && mode == MODE_EDGE && lean_edges && edge_simple(m) && LSX_CW * m - 31 - TBX_LAG * (TBX_T - 1) >= 2) {
                fast_steps(std::integral_constant<int, 2>{});       // end of the rows only
            } else if (mode == MODE_FAST) {
                // short last group (K not a multiple of TBX_T): same thing with the sub-step count tested
                for (int s = LSX_CW * m; s < s_end; ++s) {
#pragma unroll
                    for (int t = 0; t < TBX_T; ++t) {
                        if (TSEL >= 0 && t != TSEL) continue;
                        if (t < nsub) {
                            const uint32_t o = ((uint32_t)(s - lane - TBX_LAG * t) & TBX_CMASK) << 2;
                            const uint32_t o1 = (o + 4u) & TBX_OMASK, om1 = (o - 4u) & TBX_OMASK;
                            float up = __shfl_up_sync(0xffffffffu, cur[t], 1);
                            const float right = lds_f32(xs_row[t] + o1);
                            const float down = lds_f32(down_row[t] + o);
                            const float x0v = lds_f32(x0_row[t] + o);
                            if (first[t]) up = lds_f32(top_base[t] + o);
                            const float newv = gs_update(x0v, right, cur[t], down, up, a, c_recip);
                            float F = cur[t];
                            if (ORIENT == EQ_ADJUST_COLUMN) {
                                const float below = lds_f32(xs_row[t] + TBX_ROWB + om1);
                                F = col_top[t] ? -prev_up[t] : (col_bot[t] ? -below : F);
                            }
                            if (in_row[t]) sts_f32(xs_row[t] + om1, F);
                            if (ORIENT == EQ_PASSIVE) {
                                if (frame_top[t]) sts_f32(xs_row[t] - TBX_ROWB + om1, cur[t]);
                                if (frame_bot[t]) sts_f32(xs_row[t] + TBX_ROWB + om1, cur[t]);
                            }
                            if (lane == 31) sts_f32(raw_out + (uint32_t)t * TBX_ROWB + o, newv);
                            prev2[t] = cur[t];
                            prev_up[t] = up;
                            cur[t] = newv;
                            __syncwarp();
                        }
                    }
                }
            } else if (mode == MODE_EDGE) {
                general_steps(std::true_type{});
            } else {
                general_steps(std::false_type{});
            }
            TBX_TRACE(3, m);
            if (TSEL >= 0) {
                // the compute warps meet between macro steps (an mbarrier rather than bar.sync: its wait has the
                // watchdog and sees the error flag, so a failing warp can never leave the other one stuck)
                __syncwarp();
                if (lane == 0) mbar_arrive(sbase + TBX_AB_OFF);
                if (!lsx_wait_bar(sbase + TBX_AB_OFF, (uint32_t)(m & 1), p.error, lane)) return false;
            }
            if (TSEL <= 0 && m >= TBX_BACK && m - TBX_BACK < NC && lane == 0) mbar_arrive(bar_done(m - TBX_BACK));
        }
        if (TSEL <= 0 && lane == 0)
            for (int q = max(M - TBX_BACK, 0); q < NC; ++q) mbar_arrive(bar_done(q));
        return true;
    }
};

__global__ void __launch_bounds__(TBX_THREADS, 5) k_linsolve_tb(const TbxParams p) {
    if (p.run_if && *p.run_if == 0u) return;
    EQ_DYN_SMEM(tbx_smem_raw);
    const uint32_t sbase = smem_u32(tbx_smem_raw);
    const int total = p.njobs * p.nprob;
    const int lane = (int)threadIdx.x & 31;
    if (threadIdx.x == 0) sts_u32(sbase + TBX_MISC_OFF + 8u, (p.rotate_roles && !TBX_SPLIT) ? eq_cta_slot_rotation() : 0u);
    __syncthreads();
    // role 0 compute, 1 loader, 2 storer, 3 publisher (uniform per warp: taken through a shuffle so that
    // the compiler keeps the dispatch branch-uniform)
    // (TBX_SPLIT: warps 4.. are the compute warps of sub-steps 1..; 5-warp CTAs already land on rotating schedulers)
    const int wraw = (int)threadIdx.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, wraw < 4 ? ((wraw - (int)lds_u32(sbase + TBX_MISC_OFF + 8u)) & 3) : wraw, 0);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = (ld_volatile_s32(p.error) != 0) ? 0xffffffffu : atomicAdd(p.ticket, 1u);
            sts_u32(sbase + TBX_MISC_OFF, t);
            sts_u32(sbase + TBX_MISC_OFF + 4u, 0u);
            // (re-initialised per job: only sound while a job completes an even number of phases per slot, see single_shot)
            for (int i = 0; i < TBX_SLOTS; ++i) {
                mbar_init(sbase + TBX_BAR_OFF + (uint32_t)i * 16u, 32u);
                mbar_init(sbase + TBX_BAR_OFF + (uint32_t)(TBX_SLOTS + i) * 16u, 1u);
                mbar_init(sbase + TBX_BAR_OFF + (uint32_t)(2 * TBX_SLOTS + i) * 16u, 1u);
            }
            mbar_init(sbase + TBX_AB_OFF, (uint32_t)TBX_CWARPS);
        }
        __syncthreads();
        const unsigned t = lds_u32(sbase + TBX_MISC_OFF);
        if (t >= (unsigned)total) break;
        const int pi = (int)(t % (unsigned)p.nprob);
        const uint32_t jb = p.jobs[t / (unsigned)p.nprob];
        const int g = (int)(jb >> 16), b = (int)(jb & 0xffffu);
        const TbxProblem &pr = p.prob[pi];
        if (p.jobtimes && threadIdx.x == 0 && pi == 0) {
            p.jobtimes[4 * ((size_t)g * p.NBP + b)] = lsx_gtime();
#ifndef EQ_HOST_EMU
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.jobtimes[4 * ((size_t)g * p.NBP + b) + 3] = smid;
#endif
        }
#define TBX_DISPATCH(O)                                           \
    {                                                             \
        const TbxJob<O> job(p, pr, sbase, b, g, lane);            \
        if (warp == 0) job.template run_compute<TBX_SPLIT ? 0 : -1>();   \
        else if (warp == 1) job.run_loader();                     \
        else if (warp == 2) job.run_storer();                     \
        else if (warp == 3) job.run_publisher();                  \
        else job.template run_compute<TBX_SPLIT ? 1 : -1>();      \
    }
        if (pr.orient == EQ_ADJUST_ROW) TBX_DISPATCH(EQ_ADJUST_ROW)
        else if (pr.orient == EQ_ADJUST_COLUMN) TBX_DISPATCH(EQ_ADJUST_COLUMN)
        else TBX_DISPATCH(EQ_PASSIVE)
#undef TBX_DISPATCH
        __syncthreads();
        if (p.jobtimes && threadIdx.x == 0 && pi == 0) p.jobtimes[4 * ((size_t)g * p.NBP + b) + 2] = lsx_gtime();
        if (p.single_shot) break;
    }
}
