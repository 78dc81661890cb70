// k_linsolve_wf.cuh -- bit-exact lexicographic Gauss-Seidel (fluid.rs:301-325) as a space-time
// wavefront whose temporal blocking lives in REGISTERS.
//
// Same decomposition as k_linsolve_tb.cuh (band of 32 rows x all columns x WF_T consecutive
// iterations = one job run by one compute warp, lane = row, lanes one column apart, rows of the
// band move up WF_SH = 2 rows per iteration so a job only depends on the band above in its own
// iteration group and on the band below in the previous group), but the iterations of a job talk
// to each other through shuffles instead of a shared-memory tile:
//
//   sub-step t, lane r:  row  rho = 32 b - 2 t + r,   column at step s:  c = s - r - t - 1
//   R_t = value after the sweep of iteration k0+t ("raw"), F_t = after its set_boundaries ("fixed")
//   left   R_t(c-1, rho)        own previous result
//   up     R_t(c, rho-1)        lane r-1, previous step          (lane 0: band above, "raw" stream)
//   right  F_{t-1}(c+1, rho)    lane r-2 of sub-step t-1, finalised in the previous step
//   down   F_{t-1}(c, rho+1)    lane r-1 of sub-step t-1, finalised in the previous step
//                               (lanes 0,1: lanes 30,31 of the band above, "edge" streams; t = 0: x tile)
//   x0(c, rho)                  shared-memory tile
// A cell is finalised one step after it was computed: F_t(c-1) is R_t(c-1) unless the cell mirrors
// a neighbour (fluid.rs:133-189), and every possible source is at hand then: R_t(c-2) and R_t(c)
// in the lane's registers, R_t(c-1, rho-1) is the `up` of the previous step, R_t(c-1, rho+1) is what
// lane r+1 computes in this very step (shfl_down).  Frame rows and columns travel through the same
// pipeline as pass-through cells (their "update" is the identity), and the Passive frame copies
// (fluid.rs:179-187, quirk Q6) are the same four mirror codes with sign +.
//
// Why: ncu + stage isolation of k_linsolve_tb (profiles/r02_wf_notes.md) showed its single compute
// warp to be the limit -- 62 instructions per step for two iterations, every shared-memory access an
// `asm volatile` that pins the instruction order, so a step took 132 cycles alone and 211 with five
// jobs per SM.  Here a step of FOUR iterations is ~60 instructions that the compiler is free to
// interleave (plain C++ accesses, operands of step s+1 fetched before the stores of step s).
//
// Shared memory (per job): the tiles are stored DE-SKEWED -- row i of a tile is rotated by rot(i)
// positions so that at step s every lane of every sub-step reads position  s + const  of its row:
// all addresses in the unrolled step loop are [register + immediate], and an odd row stride makes
// the 32 lanes of a step hit 32 different banks.  The loader rotates while it stages (4-byte
// cp.async; it is ordered after an acquire of the producers' progress flags).
//
//   XIN  33 rows  x  WF_RX   x(t = 0 operands) rows 32b .. 32b+32                  rot(i) = i
//   X0   32+SK rows x WF_RX(+16 mirrored)  rows 32b-SK .. 32b+31                   rot(i) = i - SK + 1
//   IE   T x 3 x WF_RI       raw_t (lane 0), edge A_t (lane 30), edge B_t (lane 31) of the band above
//   OT   32 rows x WF_RO     F of the job's last sub-step (what goes back to x)     rot(r) = r + tl + 2
//   OE   T x 2 x WF_RO float2  {F_t, R_t} of lanes 30, 31 for the band below
//   CD   33+SK rows x 128 B  per-cell fix-up codes, not rotated (general steps only)
// Roles: compute, loader, storer, publisher (one warp each, see k_linsolve_exact.cuh for why).
#pragma once
#include "k_linsolve_exact.cuh"
#include <type_traits>

#ifndef WF_T
#define WF_T 4                       // iterations per job
#endif
#define WF_SH 2                      // rows the band moves up per iteration
#define WF_LG 1                      // columns between consecutive sub-steps (WF_SH + WF_LG = 3)
#define WF_CW 16                     // chunk = macro step = 16 columns / steps
#ifndef WF_UN
#define WF_UN 4                      // steps per straight-line block of the step loops
#endif
#ifndef WF_RX
#define WF_RX 80                     // ring positions of the x and x0 tiles (multiple of 16)
#endif
#ifndef WF_RI
#define WF_RI 64                     // ring positions of the in-edge streams
#endif
#ifndef WF_RO
#define WF_RO 80                     // ring positions of the output tile and the out-edge streams
#endif
#define WF_SK (WF_SH * (WF_T - 1))   // rows the band has moved up at its last sub-step
#define WF_XROWS 33
#define WF_X0ROWS (32 + WF_SK)
#define WF_CROWS (33 + WF_SK)
#define WF_XS (WF_RX + 1)            // row strides in floats: odd, so one position of 32 rows = 32 banks
#define WF_X0S (WF_RX + 17)          // + 16 mirrored positions (a sub-step reads up to 15 past its wrapped base)
#define WF_OS (WF_RO + 1)
#define WF_OES (WF_RO + 1)           // float2 per out-edge row: + 1 so lanes 30 / 31 of one STS.64 use different banks
#define WF_CDS 128                   // bytes per code row (power of two, not rotated)
#define WF_BACK 3                    // after macro step q+3 chunk q of the job's outputs is complete
#define WF_LEAD ((WF_RX - 32 - 3 * (WF_T - 1)) / 16)   // the loader may be this many waves ahead of the oldest unfinished macro step
#define WF_NBAR 8
static_assert(WF_T >= 2 && WF_T <= 8, "rows of one band must stay inside two bands of the previous group and inside the row padding");
static_assert(WF_RX % 16 == 0 && WF_RI % 16 == 0 && WF_RO % 16 == 0, "rings hold whole chunks");
static_assert((WF_XS & 1) && (WF_X0S & 1) && (WF_OS & 1), "odd strides");
static_assert(WF_LEAD >= 2 && WF_LEAD < WF_NBAR - 2, "prefetch depth");
static_assert(WF_RI > 16 * WF_LEAD + 16 + WF_T + 1, "in-edge ring vs loader lead");
static_assert(WF_SH * (WF_T - 1) + 32 < EQ_ROW_PAD + 1, "the last band's tile rows stay inside the allocation");

#define WF_XIN_OFF 0u
#define WF_X0_OFF (WF_XIN_OFF + WF_XROWS * WF_XS * 4u)
#define WF_IE_OFF (WF_X0_OFF + WF_X0ROWS * WF_X0S * 4u)
#define WF_OT_OFF (WF_IE_OFF + WF_T * 3u * WF_RI * 4u)
#define WF_OE_OFF ((WF_OT_OFF + 32u * WF_OS * 4u + 7u) & ~7u)
#define WF_CD_OFF ((WF_OE_OFF + WF_T * 2u * WF_OES * 8u + 15u) & ~15u)
#define WF_BAR_OFF (WF_CD_OFF + WF_CROWS * WF_CDS)
#define WF_MISC_OFF (WF_BAR_OFF + 2u * WF_NBAR * 16u)
#define WF_SMEM_BYTES (WF_MISC_OFF + 32u)
#define WF_THREADS 128

struct WfProblem {
    float *x;
    const float *x0;
    float *raw;          // [T][NBP][P]       R_t of lane 31 of band b-1, read by band b
    float *edge;         // [T-1][NBP][2][P]  F_t of lanes 30, 31 of band b-1, read by band b at sub-step t+1
    unsigned *progress;  // [G][NBP] chunks of the job's outputs that are visible
    float a, c_recip;
    int orient;
};

struct WfParams {
    WfProblem prob[2];
    int nprob;
    const uint8_t *codes;
    const uint8_t *row_fluid;    // [N] row j holds a NoWall cell (quirk Q6)
    const uint8_t *flags;        // [3 orientations][NBP][NC]: rows 32b-SK-1 .. 32b+31 hold a mirror code of that
                                 // orientation in this chunk that the role-coded loop does not know (k_build_wf_flags)
    const uint32_t *jobs;        // [G*NBP] (g << 16 | b) in wavefront order w = b + 2g
    int njobs;
    int N, P, K, G, NBP, NC;
    unsigned *ticket;
    int *error;
    int pub_batch;
    int rotate_roles;
    const unsigned *run_if;      // not null: return at once when *run_if == 0 (see LsxParams::run_if)
    int force_general;           // EQ_WF_GENERAL=1: every macro step takes the general loop (tests)
    int debug_nodeps;            // -DEQ_DEBUG_KNOBS builds only
    unsigned long long *jobtimes; // EQ_WF_JOBTIMES=file: [njobs][4] ns: ticket taken, first flags seen, first wave landed, compute done
    unsigned long long *trace;   // EQ_WF_TRACE=1: [64 bands][16 events] %globaltimer of group 0 (see eq_api.cu)
    unsigned *dbg;               // EQ_WF_DEBUG=1: [grid][32] last (job, index, wait kind) of every role, dumped when the watchdog fires
};
// role r (0 compute, 1 loader, 2 storer, 3 publisher): words 8r .. 8r+3 = job, index, what it is waiting for, jobs done
#define WF_JT(ev) do { if (p.jobtimes && lane == 0) p.jobtimes[4 * ((size_t)g * NBP + b) + (ev)] = lsx_gtime(); } while (0)
#define WF_TRACE(ev) do { if (p.trace && lane == 0 && g == 0 && b < 64) p.trace[b * 16 + (ev)] = lsx_gtime(); } while (0)
#define WF_DBG(role, idx, kind) do { if (p.dbg && lane == 0) { unsigned *d_ = p.dbg + (size_t)blockIdx.x * 32 + 8 * (role); d_[0] = ((unsigned)g << 16) | (unsigned)b; d_[1] = (unsigned)(idx); d_[2] = (unsigned)(kind); } } while (0)

template <int ORIENT>
__device__ __forceinline__ unsigned wf_decode(unsigned byte) {
    if (ORIENT == EQ_ADJUST_ROW) return byte & 3u;                                  // 1 = L, 2 = R
    if (ORIENT == EQ_ADJUST_COLUMN) { const unsigned c = (byte >> 2) & 3u; return c ? c + 2u : 0u; }   // 1 -> U, 2 -> D
    return (byte >> EQ_CODE_PASSIVE_SHIFT) & 7u;
}

template <int ORIENT>
struct WfJob {
    const WfParams &p;
    const WfProblem &pr;
    unsigned char *sm;
    uint32_t sbase;
    int b, g, lane;
    int N, P, NC, NBP, M, MP, nsub, k0, tl, jtop;

    __device__ __forceinline__ WfJob(const WfParams &p_, const WfProblem &pr_, unsigned char *sm_, uint32_t sbase_, int b_, int g_, int lane_)
        : p(p_), pr(pr_), sm(sm_), sbase(sbase_), b(b_), g(g_), lane(lane_) {
        N = p.N; P = p.P; NC = p.NC; NBP = p.NBP;
        k0 = g * WF_T;
        nsub = min(WF_T, p.K - k0);
        tl = nsub - 1;
        jtop = 32 * b;                               // row of lane 0 at sub-step 0
        M = NC + WF_BACK;                            // steps 0 .. P + 32 + LG(T-1) in macro steps of 16
        // The mbarriers are initialised ONCE per CTA: on B200 re-initialising them between jobs (with or without
        // mbarrier.inval) left the old phase in place, so a job must complete an EVEN number of phases on every slot to
        // hand the barriers to the next job of the CTA with the parity they started with: M is padded with empty macro
        // steps (the loader only arrives, the compute warp only waits and arrives) to a multiple of 2 * WF_NBAR.
        MP = (M + 2 * WF_NBAR - 1) / (2 * WF_NBAR) * (2 * WF_NBAR);
    }
    __device__ __forceinline__ uint32_t bar_full(int q) const { return sbase + WF_BAR_OFF + (uint32_t)(q % WF_NBAR) * 16u; }
    __device__ __forceinline__ uint32_t bar_mdone(int m) const { return sbase + WF_BAR_OFF + (uint32_t)(WF_NBAR + m % WF_NBAR) * 16u; }
    __device__ __forceinline__ uint32_t use_parity(int q) const { return (uint32_t)((q / WF_NBAR) & 1); }
    __device__ __forceinline__ float *xin() const { return reinterpret_cast<float *>(sm + WF_XIN_OFF); }
    __device__ __forceinline__ float *x0t() const { return reinterpret_cast<float *>(sm + WF_X0_OFF); }
    __device__ __forceinline__ float *ie() const { return reinterpret_cast<float *>(sm + WF_IE_OFF); }
    __device__ __forceinline__ float *ot() const { return reinterpret_cast<float *>(sm + WF_OT_OFF); }
    __device__ __forceinline__ float2 *oe() const { return reinterpret_cast<float2 *>(sm + WF_OE_OFF); }
    __device__ __forceinline__ unsigned char *cd() const { return sm + WF_CD_OFF; }
    // does any row of any active sub-step of this band lie outside the interior 1 .. N-2?
    __device__ __forceinline__ bool has_special_rows() const { return jtop - WF_SH * tl < 1 || jtop + 31 > N - 2; }

    // ------------------------------------------------------------------ LOADER warp
    // Wave q stages, 16 columns per row: x and x0 chunk q - (rot / 16) of every tile row (rows further down the band
    // are read later, so they are staged later: the live window of every row is the same), the in-edge streams and
    // the codes of chunk q.  Elements land rotated (header).
    __device__ __forceinline__ bool run_loader() const {
        const float *__restrict__ x = pr.x;
        const float *__restrict__ x0 = pr.x0;
        const unsigned *flag_prev = (g > 0 && !p.debug_nodeps) ? pr.progress + (size_t)(g - 1) * NBP + min(b + 1, NBP - 1) : nullptr;
        const unsigned *flag_above = (b > 0 && !p.debug_nodeps) ? pr.progress + (size_t)g * NBP + (b - 1) : nullptr;
        const int col = lane & 15, half = lane >> 4;
        const uint32_t xin_a = sbase + WF_XIN_OFF, x0_a = sbase + WF_X0_OFF, ie_a = sbase + WF_IE_OFF, cd_a = sbase + WF_CD_OFF;
        int w0 = 0, wi = 0;                                       // (16 q) mod RX, mod RI
        for (int q = 0; q < MP; ++q) {
            if (q >= M) {                                         // padding wave: keep the phase count of the slot even
                if (!lsx_wait_bar(bar_mdone(q - WF_LEAD - 1), use_parity(q - WF_LEAD - 1), p.error, lane)) return false;
                cp_async_mbar_arrive_noinc(bar_full(q));
                continue;
            }
            if (q + LSX_PF < NC) prefetch_l2(x0 + (size_t)min(jtop + lane, N - 1) * P + WF_CW * (q + LSX_PF));
            // ring space: macro steps <= q - WF_LEAD - 1 must be finished
            WF_DBG(1, q, 1);
            if (q > WF_LEAD && !lsx_wait_bar(bar_mdone(q - WF_LEAD - 1), use_parity(q - WF_LEAD - 1), p.error, lane)) return false;
            const unsigned need = (unsigned)min(q + 1, NC);
            WF_DBG(1, q, 2);
            if (q == 0) WF_TRACE(0);
            if (!lsx_wait_flags(flag_prev, need, false, flag_above, need, false, p.error, lane)) return false;
            if (q == 0) WF_TRACE(1);
            if (q == 0) WF_JT(1);
            if (q == 4) WF_TRACE(9);
            // (16 (q - g)) mod RX for the three row groups; every ring index below is base + small offset, one wrap
            const int w1 = w0 >= 16 ? w0 - 16 : w0 - 16 + WF_RX, w2 = w1 >= 16 ? w1 - 16 : w1 - 16 + WF_RX;
            // x tile: row i <-> global row jtop + i, rot(i) = i, chunk q - i/16
#pragma unroll
            for (int it = 0; it < (WF_XROWS + 1) / 2; ++it) {
                const int i = 2 * it + half, ci = q - (it >> 3);          // (2 it + half) / 16 = it / 8
                if (i < WF_XROWS && ci >= 0 && ci < NC) {
                    int pos = ((it >> 3) == 0 ? w0 : ((it >> 3) == 1 ? w1 : w2)) + col + i;
                    pos -= pos >= WF_RX ? WF_RX : 0;
                    cp_async_4s(xin_a + (uint32_t)(i * WF_XS + pos) * 4u, x + (size_t)(jtop + i) * P + WF_CW * ci + col);
                }
            }
            // x0 tile: row i <-> global row jtop - SK + i, rot(i) = i - SK + 1, chunk q - max(i - SK, 0)/16
#pragma unroll
            for (int it = 0; it < (WF_X0ROWS + 1) / 2; ++it) {
                const int i = 2 * it + half;
                const int gi = max(i - WF_SK, 0) >> 4, ci = q - gi;
                const int j = jtop - WF_SK + i;
                if (i < WF_X0ROWS && ci >= 0 && ci < NC && j >= 0) {
                    int pos = (gi == 0 ? w0 : (gi == 1 ? w1 : w2)) + col + i - WF_SK + 1;
                    pos += pos < 0 ? WF_RX : 0;
                    pos -= pos >= WF_RX ? WF_RX : 0;
                    const float *src = x0 + (size_t)j * P + WF_CW * ci + col;
                    cp_async_4s(x0_a + (uint32_t)(i * WF_X0S + pos) * 4u, src);
                    if (pos < 16) cp_async_4s(x0_a + (uint32_t)(i * WF_X0S + WF_RX + pos) * 4u, src);
                }
            }
            if (q < NC) {
                const int c = WF_CW * q + col;
                if (b > 0) {
                    // raw_t -> IR_t (rot LG t + 1), edge A_t -> IA_t (rot LG (t+1)), edge B_t -> IB_t (rot LG (t+1) + 1)
#pragma unroll
                    for (int it = 0; it < (3 * WF_T - 2 + 1) / 2; ++it) {
                        const int id = 2 * it + half;
                        if (id < nsub) {
                            const int t = id;
                            int pos = wi + col + WF_LG * t + 1;
                            pos -= pos >= WF_RI ? WF_RI : 0;
                            cp_async_4s(ie_a + (uint32_t)((t * 3 + 0) * WF_RI + pos) * 4u, pr.raw + ((size_t)t * NBP + b) * P + c);
                        } else if (id >= WF_T && id < 3 * WF_T - 2 && id - WF_T < 2 * (nsub - 1)) {
                            const int t = (id - WF_T) >> 1, w = (id - WF_T) & 1;
                            int pos = wi + col + WF_LG * (t + 1) + w;
                            pos -= pos >= WF_RI ? WF_RI : 0;
                            cp_async_4s(ie_a + (uint32_t)((t * 3 + 1 + w) * WF_RI + pos) * 4u,
                                        pr.edge + (((size_t)t * NBP + b) * 2 + w) * P + c);
                        }
                    }
                }
                // codes: row i <-> global row jtop - SK - 1 + i, 16 bytes per row and chunk, not rotated
#pragma unroll 1
                for (int i = lane; i < WF_CROWS; i += 32) {
                    const int j = jtop - WF_SK - 1 + i;
                    if (j >= 0 && j < N)
                        cp_async_16s(cd_a + (uint32_t)(i * WF_CDS + ((WF_CW * q) & (WF_CDS - 1))), p.codes + (size_t)j * P + WF_CW * q);
                }
            }
            w0 += WF_CW; w0 -= w0 >= WF_RX ? WF_RX : 0;
            wi += WF_CW; wi -= wi >= WF_RI ? WF_RI : 0;
            cp_async_mbar_arrive_noinc(bar_full(q));
            if (q == 0) WF_TRACE(2);
            WF_DBG(1, q, 3);
        }
        WF_DBG(1, M, 9);
        return true;
    }

    // ------------------------------------------------------------------ STORER warp
    __device__ __forceinline__ bool run_storer() const {
        float *__restrict__ x = pr.x;
        const float *otile = ot();
        const float2 *oedge = oe();
        const int col = lane & 15, half = lane >> 4;
        const bool has_below = (b + 1 < NBP);
        const int jl = jtop - WF_SH * tl;                         // row of lane 0 at the last sub-step
        for (int q = 0; q < NC; ++q) {
            WF_DBG(2, q, 1);
            if (!lsx_wait_bar(bar_mdone(q + WF_BACK), use_parity(q + WF_BACK), p.error, lane)) return false;
            WF_DBG(2, q, 2);
            const int c = WF_CW * q + col;
#pragma unroll 4
            for (int r = half; r < 32; r += 2) {
                const int j = jl + r;
                if (j >= 0 && j <= N - 1 && c < N) x[(size_t)j * P + c] = otile[r * WF_OS + (c + r + WF_LG * tl + 2) % WF_RO];
            }
            if (has_below) {
#pragma unroll 1
                for (int id = half; id < 3 * WF_T - 2; id += 2) {
                    if (id < nsub) {                                  // R_t of lane 31
                        const int t = id;
                        pr.raw[((size_t)t * NBP + b + 1) * P + c] = oedge[(t * 2 + 1) * WF_OES + (c + 32 + WF_LG * t) % WF_RO].y;
                    } else if (id >= WF_T && id - WF_T < 2 * (nsub - 1)) {   // F_t of lanes 30, 31
                        const int t = (id - WF_T) >> 1, w = (id - WF_T) & 1;
                        pr.edge[(((size_t)t * NBP + b + 1) * 2 + w) * P + c] = oedge[(t * 2 + w) * WF_OES + (c + 32 + w + WF_LG * t) % WF_RO].x;
                    }
                }
            }
            __syncwarp();
            if (q == 0) WF_TRACE(7);
            if (lane == 0) sts_release_cta_u32(sbase + WF_MISC_OFF + 4u, (uint32_t)q + 1u);
        }
        WF_DBG(2, NC, 9);
        return true;
    }

    // ------------------------------------------------------------------ PUBLISHER warp
    __device__ __forceinline__ bool run_publisher() const {
        unsigned *my_flag = pr.progress + (size_t)g * NBP + b;
        const uint32_t cnt = sbase + WF_MISC_OFF + 4u;
        int q = 0, ok = 1;
        while (q < NC && ok) {
            WF_DBG(3, q, 1);
            if (lane == 0) {
                unsigned spins = 0;
                unsigned long long t0 = 0;
                int have;
                // the first chunk goes out at once (it is what lets the band below start), then in batches
                while ((have = (int)lds_acquire_cta_u32(cnt)) < min(q == 0 ? 1 : q + p.pub_batch, NC)) {
                    __nanosleep(64);
                    if ((++spins & 1023u) == 0) {
                        if (lsx_expired(t0, spins)) { *p.error = 3; ok = 0; break; }
                        if (ld_volatile_s32(p.error) != 0) { ok = 0; break; }
                    }
                }
                if (ok) {
                    const bool first = (q == 0);
                    q = have;
                    st_release_u32(my_flag, (unsigned)q);
                    if (first) WF_TRACE(8);
                }
            }
            q = __shfl_sync(0xffffffffu, q, 0);
            ok = __shfl_sync(0xffffffffu, ok, 0);
        }
        WF_DBG(3, q, 9);
        return ok != 0;
    }

    // the storer's count of chunks whose stores are issued (shared word, release / acquire at CTA scope)
    __device__ __forceinline__ bool wait_stored(int need) const {
        const uint32_t cnt = sbase + WF_MISC_OFF + 4u;
        bool ok = true;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while ((int)lds_acquire_cta_u32(cnt) < need) {            // (a plain volatile load here never saw the counter move on the GPU)
            if ((++spins & 255u) == 0) {
                if (lsx_expired(t0, spins)) {
                    if (lane == 0) *p.error = 4;
                    ok = false;
                    break;
                }
                if (ld_volatile_s32(p.error) != 0) {
                    ok = false;
                    break;
                }
            }
        }
        return __all_sync(0xffffffffu, ok) != 0;
    }

    // ------------------------------------------------------------------ COMPUTE warp
    // MODE_FAST    every row of every sub-step is interior, every lane on an interior column, no mirror code in reach: F = R.
    // MODE_ROLE    like FAST, but the band holds frame rows / rows outside the grid: pass-through rows and the mirror codes
    //              every interior column of such a band has (Passive: row 0 takes row 1, row N-1 takes row N-2;
    //              AdjustColumn: row 1 takes -row 0, row N-2 takes -row N-1) as per-lane masks.
    // MODE_EDGE    MODE_ROLE at the start / end of the rows: column range tests and the mirror codes every row has there
    //              (AdjustRow: columns 1 and N-2; Passive: the frame columns of rows that hold a NoWall cell, quirk Q6).
    // MODE_GENERAL column range tests and per-cell codes from the staged code tile (obstacles).
    // Lanes outside the grid (rows < 0 or > N-1, columns < 0 or > N-1) are NOT masked: they compute garbage that never
    // reaches a cell of the grid (a frame cell passes its own old value through, every other input of a grid cell is a
    // grid cell) and that the storer never writes back.  Per-lane conditions are all-ones / zero masks combined with one
    // LOP3 per select, not predicates: a step of four sub-steps would need a dozen of them.
    enum { MODE_FAST = 0, MODE_ROLE = 1, MODE_GENERAL = 2, MODE_EDGE = 3 };

    static __device__ __forceinline__ float sel(unsigned m, float a, float b) {          // m ? a : b
        return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
    }
    static __device__ __forceinline__ float sel_sgn(unsigned m, unsigned sgn, float a, float b) {   // m ? (a with its sign flipped by sgn) : b
        return __uint_as_float(((__float_as_uint(a) ^ sgn) & m) | (__float_as_uint(b) & ~m));
    }

    // a mask the compiler cannot trace back to the comparison it came from (it would turn every select into
    // predicate logic again: P2R / SEL chains instead of one LOP3)
    static __device__ __forceinline__ unsigned opaque(bool c) {
        unsigned m = c ? ~0u : 0u;
#ifndef EQ_HOST_EMU
        asm volatile("mov.b32 %0, %0;" : "+r"(m));
#endif
        return m;
    }

    __device__ __forceinline__ bool run_compute() const {
        const float a = pr.a, c_recip = pr.c_recip;
        const int lm1 = (lane + 31) & 31, lm2 = (lane + 30) & 31;
        const bool is_edge_lane = (lane == 0) | (lane >= 30);
        const bool hi_lane = lane >= 30;
        const unsigned m_l0 = opaque(lane == 0), m_hi = opaque(lane >= 30), m_l31 = opaque(lane == 31);
        constexpr unsigned sgn = (ORIENT != EQ_PASSIVE) ? 0x80000000u : 0u;   // set_boundaries mirrors with a minus, Passive copies
        // per sub-step constants
        int row[WF_T];
        unsigned m_inr[WF_T], m_u[WF_T], m_d[WF_T], m_rowf[WF_T];
        bool patch_gen[WF_T];
        bool any_patch_role = false;
#pragma unroll
        for (int t = 0; t < WF_T; ++t) {
            row[t] = jtop - WF_SH * t + lane;
            const bool on = (t < nsub);
            const bool inr = on && row[t] >= 1 && row[t] <= N - 2;
            m_inr[t] = opaque(inr);
            bool u = false, d = false;
            if (ORIENT == EQ_PASSIVE) { d = on && row[t] == 0; u = on && row[t] == N - 1; }
            if (ORIENT == EQ_ADJUST_COLUMN) { u = on && row[t] == 1; d = on && row[t] == N - 2; }
            m_u[t] = opaque(u);
            m_d[t] = opaque(d && lane < 31);                          // lane 31: the band below finishes the cell
            m_rowf[t] = opaque(ORIENT == EQ_PASSIVE && inr && p.row_fluid[row[t]] != 0);
            // lane 0 of a lower band finishes the DOWN mirror of the last row of the band above
            patch_gen[t] = (ORIENT == EQ_ADJUST_COLUMN) && lane == 0 && b > 0 && on && row[t] >= 1 && row[t] <= N - 1;
            // the role loops do not patch: a band whose lane 0 can be the frame row N-1 (odd N only) takes the general loop
            any_patch_role = any_patch_role || ((ORIENT == EQ_ADJUST_COLUMN) && b > 0 && on && jtop - WF_SH * t == N - 1);
        }
        // per-sub-step rows of the tiles are constant offsets from the lane's row at sub-step 0
        const float *x0p0 = x0t() + (WF_SH * (WF_T - 1) + lane) * WF_X0S;          // sub-step t: - t * SH * X0S
        const float *ep0 = ie() + (lane == 0 ? 0 : (lane == 30 ? 1 : 2)) * WF_RI;  // sub-step t: + t * 3 * RI
        float2 *oep0 = oe() + (lane & 1) * WF_OES;                                  // sub-step t: + t * 2 * OES
        const unsigned char *crow0 = cd() + (WF_SK + lane + 1) * WF_CDS;            // sub-step t: - t * SH * CDS
        const uint32_t ib0_a = sbase + WF_IE_OFF + 2u * WF_RI * 4u;                 // edge stream B_t: + t * 3 * RI
        const float *xr = xin() + lane * WF_XS, *xd = xin() + (lane + 1) * WF_XS;
        float *otp = ot() + lane * WF_OS;
        float *xg = pr.x;

        // pipeline state: cur = R_t(c-1), h2 = R_t(c-2), fh = F_t(c-2), pup = the `up` of the previous step, prgt = F_{t-1}(c, rho)
        float cur[WF_T], fh[WF_T], h2[WF_T], pup[WF_T], prgt[WF_T];
#pragma unroll
        for (int t = 0; t < WF_T; ++t) cur[t] = fh[t] = h2[t] = pup[t] = prgt[t] = 0.f;

        int ox = 0, oi = 0, oo = 0;                               // (16 m) mod ring
        int o0[WF_T];
#pragma unroll
        for (int t = 0; t < WF_T; ++t) o0[t] = (WF_RX * 4 - 3 * t) % WF_RX;

        // ---- one step; `i` = step inside the macro step (compile-time in the unrolled loops), cl = 16 m - lane.
        // Operands (rgt0, dwn0, x0v[], ev[]) are passed in: the loops fetch those of step i+1 before the stores of step i
        // (the compiler cannot prove that the output rings do not alias the tiles).
        auto step = [&](auto mode_c, int cl, int i, float rgt0, float dwn0, const float (&x0v)[WF_T], const float (&ev)[WF_T]) {
            constexpr int MODE = decltype(mode_c)::value;
            constexpr bool RANGED = (MODE == MODE_EDGE || MODE == MODE_GENERAL);
            float up[WF_T], rgt[WF_T], dwn[WF_T];
#pragma unroll
            for (int t = 0; t < WF_T; ++t) {
                up[t] = sel(m_l0, ev[t], __shfl_sync(0xffffffffu, cur[t], lm1));
                if (t == 0) {
                    rgt[t] = rgt0;
                    dwn[t] = dwn0;
                } else {
                    // lanes 30, 31 lend their slot in the rotation to the edge rows of the band above (their own F goes to
                    // the band below through OE); lane 30's F is still lane 31's `down`
                    rgt[t] = __shfl_sync(0xffffffffu, sel(m_hi, ev[t - 1], fh[t - 1]), lm2);
                    dwn[t] = __shfl_sync(0xffffffffu, sel(m_l31, ev[t - 1], fh[t - 1]), lm1);
                }
            }
#pragma unroll
            for (int t = 0; t < WF_T; ++t) {
                const float gsv = gs_update(x0v[t], rgt[t], cur[t], dwn[t], up[t], a, c_recip);
                float nv = gsv, F = cur[t];
                if (MODE != MODE_FAST) {
                    const int c = cl + (i - WF_LG * t - 1), cf = c - 1;
                    bool coli = true, cfi = true;                          // c, cf in 1 .. N-2
                    if (RANGED) {
                        coli = (unsigned)(c - 1) <= (unsigned)(N - 3);
                        cfi = (unsigned)(cf - 1) <= (unsigned)(N - 3);
                    }
                    nv = sel(m_inr[t], gsv, prgt[t]);
                    if (RANGED) nv = coli ? nv : prgt[t];
                    if (MODE == MODE_GENERAL) {
                        const unsigned code = wf_decode<ORIENT>(crow0[(cf & (WF_CDS - 1)) - t * WF_SH * WF_CDS]);
                        if (ORIENT != EQ_ADJUST_COLUMN) {
                            F = (code == WF_C_L) ? __uint_as_float(__float_as_uint(h2[t]) ^ sgn) : F;
                            F = (code == WF_C_R) ? __uint_as_float(__float_as_uint(nv) ^ sgn) : F;
                        }
                        if (ORIENT != EQ_ADJUST_ROW) {
                            const float dnv = __shfl_down_sync(0xffffffffu, nv, 1);
                            F = (code == WF_C_U) ? __uint_as_float(__float_as_uint(pup[t]) ^ sgn) : F;
                            F = ((code == WF_C_D) & (lane < 31)) ? __uint_as_float(__float_as_uint(dnv) ^ sgn) : F;
                        }
                    } else {
                        if (ORIENT != EQ_ADJUST_ROW) {                     // frame rows (per-lane masks), interior columns
                            const float dnv = __shfl_down_sync(0xffffffffu, nv, 1);
                            float Fr = sel_sgn(m_u[t], sgn, pup[t], cur[t]);
                            Fr = sel_sgn(m_d[t], sgn, dnv, Fr);
                            F = cfi ? Fr : F;
                        }
                        if (RANGED) {                                      // frame columns, interior rows
                            if (ORIENT == EQ_ADJUST_ROW) {                 // (1, j) takes -x[0, j], (N-2, j) takes -x[N-1, j]
                                F = (cf == 1) ? sel_sgn(m_inr[t], sgn, h2[t], F) : F;
                                F = (cf == N - 2) ? sel_sgn(m_inr[t], sgn, nv, F) : F;
                            } else if (ORIENT == EQ_PASSIVE) {             // (0, j) takes x[1, j], (N-1, j) takes x[N-2, j]
                                F = (cf == 0) ? sel(m_rowf[t], nv, F) : F;
                                F = (cf == N - 1) ? sel(m_rowf[t], h2[t], F) : F;
                            }
                        }
                    }
                    if (ORIENT == EQ_ADJUST_COLUMN && MODE == MODE_GENERAL) {
                        // lane 0 of a lower band: the cell above (c, row-1) takes -R_t(c, row) when its code says DOWN.  It
                        // lives on in my edge stream B_t (read by lane 31 in the next step, position s+1), or -- after the
                        // job's last sub-step -- already in global x
                        const bool patch = patch_gen[t] && coli && wf_decode<ORIENT>(crow0[(c & (WF_CDS - 1)) - WF_CDS - t * WF_SH * WF_CDS]) == WF_C_D;
                        const int pi = (oi + i + 1 >= WF_RI) ? oi + i + 1 - WF_RI : oi + i + 1;
                        st_shared_f32_if(patch && t != tl, ib0_a + (uint32_t)(pi + t * 3 * WF_RI) * 4u, -nv);
                        st_global_f32_if(patch && t == tl, xg + ((ptrdiff_t)(row[t] - 1) * P + c), -nv);   // (not dereferenced when the predicate is off)
                    }
                }
                if (hi_lane) oep0[oo + i + t * 2 * WF_OES] = make_float2(F, nv);
                if ((MODE == MODE_FAST) ? (t == WF_T - 1) : (t == tl)) otp[oo + i] = F;
                h2[t] = cur[t];
                pup[t] = up[t];
                prgt[t] = rgt[t];
                fh[t] = F;
                cur[t] = nv;
            }
        };
        auto fetch_x = [&](int i, float &rgt0, float &dwn0, float (&x0v)[WF_T]) {
            rgt0 = xr[ox + i];
            dwn0 = xd[ox + i];
#pragma unroll
            for (int t = 0; t < WF_T; ++t) x0v[t] = x0p0[o0[t] + i - t * WF_SH * WF_X0S];
        };
        auto fetch_e = [&](int i, float (&ev)[WF_T]) {
#pragma unroll
            for (int t = 0; t < WF_T; ++t) ev[t] = is_edge_lane ? ep0[oi + i + t * 3 * WF_RI] : 0.f;
        };
        // A macro step = WF_CW / WF_UN blocks of WF_UN straight-line steps with operand prefetch (the loop body has to fit
        // the instruction cache of a sub-partition: fully unrolled, the 16 steps of the fast loop were 18.6 KB of SASS and
        // ncu put half of the loop's stall samples on stall_no_inst).  `blk` is a runtime value: all ring offsets are
        // registers bumped by WF_UN per block, `i` below is the step inside the block.
        // AdjustColumn in the general loop: lane 0 may patch an entry of edge stream B that lane 31 reads in the NEXT step, so
        // the edge values are fetched at the start of their own step, after a __syncwarp.
        auto macro_unrolled = [&](auto mode_c, int m) {
            constexpr bool LATE_E = (ORIENT == EQ_ADJUST_COLUMN) && (decltype(mode_c)::value == MODE_GENERAL);
            int cl = WF_CW * m - lane;
            float rgt0, dwn0, x0v[WF_T], ev[WF_T];
            fetch_x(0, rgt0, dwn0, x0v);
            fetch_e(0, ev);
#pragma unroll 1
            for (int blk = 0; blk < WF_CW / WF_UN; ++blk) {
#pragma unroll
                for (int i = 0; i < WF_UN; ++i) {
                    float rgt0n, dwn0n, x0n[WF_T], evn[WF_T];
                    fetch_x(i + 1, rgt0n, dwn0n, x0n);           // (the last one of a macro step reads one position past it: unused)
                    if (!LATE_E) fetch_e(i + 1, evn);
                    step(mode_c, cl, i, rgt0, dwn0, x0v, ev);
                    if (LATE_E) {
                        __syncwarp();
                        fetch_e(i + 1, evn);
                    }
                    rgt0 = rgt0n;
                    dwn0 = dwn0n;
#pragma unroll
                    for (int t = 0; t < WF_T; ++t) { x0v[t] = x0n[t]; ev[t] = evn[t]; }
                }
                ox += WF_UN; oi += WF_UN; oo += WF_UN; cl += WF_UN;
#pragma unroll
                for (int t = 0; t < WF_T; ++t) o0[t] += WF_UN;
            }
        };

        const uint8_t *fl = p.flags + ((size_t)ORIENT * NBP + b) * NC;
        const bool special = has_special_rows() || nsub < WF_T;
        unsigned fl_hist = 0, fl_next = fl[0] != 0 ? 1u : 0u;
        long long cyc_full = 0, cyc_pre = 0, cyc_body = 0;          // EQ_WF_TRACE: where the compute warp's time goes
        for (int m = 0; m < MP; ++m) {
            WF_DBG(0, m, 1);
            const long long c0 = p.trace ? lsx_clock() : 0;
            if (!lsx_wait_bar(bar_full(m), use_parity(m), p.error, lane)) return false;
            const long long c1 = p.trace ? lsx_clock() : 0;
            WF_DBG(0, m, 2);
            if (m == 0) WF_TRACE(3);
            if (m == 0) WF_JT(2);
            if (m == 1) WF_TRACE(4);
            if (m == 3) WF_TRACE(5);
            if (m == 4) WF_TRACE(6);
            if (m == 8) WF_TRACE(10);
            if (m == 64) WF_TRACE(11);
            if (m == 128) WF_TRACE(12);
            if (m >= M) {                                         // padding macro step
                if (lane == 0) mbar_arrive(bar_mdone(m));
                continue;
            }
            // output rings: this macro step overwrites the positions of macro step m - RO/16, whose newest data belong to
            // chunk m - RO/16 -- it must have been stored (the storer's counter; it is almost always far ahead)
            if (m >= WF_RO / WF_CW && !wait_stored(min(m - WF_RO / WF_CW + 1, NC))) return false;
            __syncwarp();      // the polling loops have lane-dependent exits as far as the compiler can tell: without an explicit
                               // convergence point it wraps every shuffle of the step loops in WARPSYNC.COLLECTIVE
            WF_DBG(0, m, 3);
            // columns touched: finalised c-1 >= 16m - 31 - LG(T-1) - 2, right operand c+1 <= 16m + 15
            const int cmin = WF_CW * m - 33 - WF_LG * (WF_T - 1), cmax = WF_CW * m + 15;
            const bool edge = cmin < 1 || cmax > N - 2;
            // codes in reach: chunks m-3 .. m (a sliding window over the chunk flags; the next flag is loaded a macro step ahead)
            fl_hist = ((fl_hist << 1) | fl_next) & ((2u << WF_BACK) - 1u);
            fl_next = fl[min(m + 1, NC - 1)] != 0 ? 1u : 0u;
            const unsigned any = fl_hist;
            const bool general = p.force_general || any != 0 || N < 8 || any_patch_role;
            const long long c2 = p.trace ? lsx_clock() : 0;
            if (general) macro_unrolled(std::integral_constant<int, MODE_GENERAL>{}, m);
            else if (edge) macro_unrolled(std::integral_constant<int, MODE_EDGE>{}, m);
            else if (special) macro_unrolled(std::integral_constant<int, MODE_ROLE>{}, m);
            else macro_unrolled(std::integral_constant<int, MODE_FAST>{}, m);
            if (ox >= WF_RX) ox -= WF_RX;                         // (the blocks of the macro step advanced the offsets by 16)
            if (oi >= WF_RI) oi -= WF_RI;
            if (oo >= WF_RO) oo -= WF_RO;
#pragma unroll
            for (int t = 0; t < WF_T; ++t) if (o0[t] >= WF_RX) o0[t] -= WF_RX;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_mdone(m));
            if (p.trace) { cyc_full += c1 - c0; cyc_pre += c2 - c1; cyc_body += lsx_clock() - c2; }
        }
        if (p.trace && lane == 0 && g == 0 && b < 64) {
            p.trace[b * 16 + 13] = (unsigned long long)cyc_full;
            p.trace[b * 16 + 14] = (unsigned long long)cyc_pre;
            p.trace[b * 16 + 15] = (unsigned long long)cyc_body;
        }
        WF_DBG(0, M, 9);
        WF_JT(3);
        return true;
    }
};

#ifndef WF_CTAS_PER_SM
#define WF_CTAS_PER_SM 4
#endif
__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM) k_linsolve_wf(const WfParams p) {
    if (p.run_if && *p.run_if == 0u) return;
    EQ_DYN_SMEM(wf_smem_raw);
    const uint32_t sbase = smem_u32(wf_smem_raw);
    const int total = p.njobs * p.nprob;
    const int lane = (int)threadIdx.x & 31;
    if (threadIdx.x == 0) {
        sts_u32(sbase + WF_MISC_OFF + 8u, p.rotate_roles ? eq_cta_slot_rotation() : 0u);
        for (int i = 0; i < WF_NBAR; ++i) {
            mbar_init(sbase + WF_BAR_OFF + (uint32_t)i * 16u, 32u);                  // full: 32 loader lanes (cp.async arrive.noinc)
            mbar_init(sbase + WF_BAR_OFF + (uint32_t)(WF_NBAR + i) * 16u, 1u);       // mdone: compute lane 0
        }
    }
    __syncthreads();
    const int wraw = (int)threadIdx.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, (wraw - (int)lds_u32(sbase + WF_MISC_OFF + 8u)) & 3, 0);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = (ld_volatile_s32(p.error) != 0) ? 0xffffffffu : atomicAdd(p.ticket, 1u);
            sts_u32(sbase + WF_MISC_OFF, t);
            sts_u32(sbase + WF_MISC_OFF + 4u, 0u);
        }
        __syncthreads();
        const unsigned t = lds_u32(sbase + WF_MISC_OFF);
        if (t >= (unsigned)total) break;
        const int pi = (int)(t % (unsigned)p.nprob);
        const uint32_t jb = p.jobs[t / (unsigned)p.nprob];
        const int g = (int)(jb >> 16), b = (int)(jb & 0xffffu);
        const WfProblem &pr = p.prob[pi];
        if (p.jobtimes && threadIdx.x == 0) p.jobtimes[4 * ((size_t)g * p.NBP + b)] = lsx_gtime();
#define WF_DISPATCH(O)                                            \
    {                                                             \
        const WfJob<O> job(p, pr, wf_smem_raw, sbase, b, g, lane); \
        if (warp == 0) job.run_compute();                         \
        else if (warp == 1) job.run_loader();                     \
        else if (warp == 2) job.run_storer();                     \
        else job.run_publisher();                                 \
    }
        if (pr.orient == EQ_ADJUST_ROW) WF_DISPATCH(EQ_ADJUST_ROW)
        else if (pr.orient == EQ_ADJUST_COLUMN) WF_DISPATCH(EQ_ADJUST_COLUMN)
        else WF_DISPATCH(EQ_PASSIVE)
#undef WF_DISPATCH
    }
}

// (band, chunk) summaries for k_linsolve_wf: does the chunk hold a mirror code of the orientation that the role-coded
// loop does not already know?  One thread per cell.
__global__ void k_build_wf_flags(const uint8_t *__restrict__ codes, uint8_t *flags, int NBP, int NC, EqLayout L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= L.N) return;
    const int N = L.N;
    const unsigned byte = codes[(size_t)j * L.P + i];
    const bool wall = (byte & EQ_CODE_WALL) != 0;
    const bool interior_col = (i >= 1 && i <= N - 2);
    bool f[3];
    // AdjustRow: columns 1 / N-2 are expected to mirror the frame (LEFT / RIGHT) in every interior row
    const bool interior_row = (j >= 1 && j <= N - 2);
    f[EQ_ADJUST_ROW] = (byte & 3u) != 0;
    if (interior_row && i == 1) f[EQ_ADJUST_ROW] = wall || (byte & 3u) != EQ_CODE_ROW_LEFT;
    else if (interior_row && i == N - 2) f[EQ_ADJUST_ROW] = wall || (byte & 3u) != EQ_CODE_ROW_RIGHT;
    // AdjustColumn: rows 1 / N-2 are expected to mirror the frame (UP / DOWN) in every interior column
    unsigned cc = (byte >> 2) & 3u;
    f[EQ_ADJUST_COLUMN] = cc != 0;
    if (interior_col && j == 1) f[EQ_ADJUST_COLUMN] = wall || (byte & 12u) != EQ_CODE_COL_UP;
    else if (interior_col && j == N - 2) f[EQ_ADJUST_COLUMN] = wall || (byte & 12u) != EQ_CODE_COL_DOWN;
    if (N < 5) f[EQ_ADJUST_COLUMN] = f[EQ_ADJUST_COLUMN] || (j >= 1 && j <= N - 2);   // rows 1 and N-2 coincide or touch
    // Passive: the frame rows are expected to copy in every interior column; frame columns only occur in edge steps
    const unsigned pc = (byte >> EQ_CODE_PASSIVE_SHIFT) & 7u;
    f[EQ_PASSIVE] = false;
    if (interior_col && j == 0) f[EQ_PASSIVE] = pc != WF_C_D;
    else if (interior_col && j == N - 1) f[EQ_PASSIVE] = pc != WF_C_U;
    // band b reads the codes of rows 32b - SK - 1 .. 32b + 31
    for (int o = 0; o < 3; ++o) {
        if (!f[o]) continue;
        for (int bb = max((j - 31 + 31) / 32, 0); bb <= min((j + WF_SK + 1) / 32, NBP - 1); ++bb)
            if (j >= 32 * bb - WF_SK - 1 && j <= 32 * bb + 31) flags[((size_t)o * NBP + bb) * NC + i / WF_CW] = 1;
    }
}
