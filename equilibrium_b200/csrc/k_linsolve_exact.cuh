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
#define LSX_CW EQ_LSX_CW                               // chunk width in columns (16): unit of staging,
                                                       // of write-back and of the progress flags
#define LSX_SLOTS (128 / LSX_CW)                       // the 128-column ring holds this many chunks
#define LSX_BACK (32 / LSX_CW)                         // after macro step m, chunk m-LSX_BACK is final
#define LSX_BAR_OFF (LSX_RAW_OFF + 128u * 4u)          // mbarriers: full[], done[], free[], 16 B each
#define LSX_MISC_OFF (LSX_BAR_OFF + 3u * LSX_SLOTS * 16u)   // [0] ticket broadcast, [1] chunks whose stores are issued
#define LSX_SMEM_BYTES (LSX_MISC_OFF + 16u)
#define LSX_THREADS 128
#define LSX_PF 24                                      // L2 prefetch distance of the loader, in chunks
#define LSX_SPIN_LIMIT (1u << 22)          // emulated build: spins before a wait gives up
#define LSX_TIMEOUT_NS 4000000000ull       // device: a wait gives up after 4 s of %globaltimer (a spin count is not a
                                           // time: jobs deep in the chain legitimately wait tens of ms before their first chunk)

struct LsxProblem {
    float *x;            // in/out, in place
    const float *x0;
    float *raw;          // [NB][P] raw stream: R_k of the last row of band b-1, read by band b
    unsigned *progress;  // [K][NB] chunks completed
    float a, c_recip;
    int orient;          // EqOrientation
    // row slabs over several GPUs: the same arrays on rank-1 / rank+1 (nullptr on one GPU)
    float *x_up;         // rank-1's x: my first band writes its first row and its DOWN patches there
    float *raw_down;     // rank+1's raw stream: my last band writes R_k of its last row there
    unsigned *prog_up, *prog_down;   // the neighbours' mirrors of my boundary bands' progress
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
    int b_lo, b_hi;              // bands [b_lo, b_hi) are mine (row slab)
    unsigned *ticket;
    int *error;
    int slack;                   // a consumer asks its producers to be this many chunks further ahead than strictly needed
    int pub_batch;               // the publisher releases progress every pub_batch chunks (and at the end): one release
                                 // costs ~4 us of fence whatever it covers; one chunk per release paces the whole chain
    int rotate_roles;            // 1: spread the roles over the warp schedulers (eq_cta_slot_rotation); EQ_LSX_ROT=0 turns it off
    int debug_nodeps;            // EQ_LSX_NODEPS=1: skip the dependency waits (WRONG results; throughput experiments only)
    const unsigned *run_if;      // not null: return at once when *run_if == 0 (the a == 0 shortcut was taken, k_a0_check)
    int single_shot;             // one job per CTA (grid = jobs): set when a job completes an ODD number of phases per mbarrier
                                 // slot -- re-initialising the barriers for a second job of the same CTA does not take on B200
                                 // (profiles/r02_wf_notes.md, section 3), the next job would wait on the wrong parity
    unsigned long long *stats;   // optional [16] cycle counters (EQ_LSX_STATS=1), see eq_api.cu
    unsigned long long *jobtimes; // optional [2 * njobs * nprob] start/end ns of every job (EQ_LSX_JOBTIMES=1)
    unsigned long long *trace;   // optional event trace of the first jobs (EQ_LSX_TRACE=1): [4 bands][8 events][128 chunks] ns
};

#ifdef EQ_HOST_EMU
static inline long long lsx_clock() { return 0; }
#else
__device__ __forceinline__ long long lsx_clock() { return clock64(); }
#endif
#ifdef EQ_HOST_EMU
static inline unsigned long long lsx_gtime() { return 0; }
#else
__device__ __forceinline__ unsigned long long lsx_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
// watchdog of the waiting loops: `t0` is the time of the first check (0 = not taken yet)
__device__ __forceinline__ bool lsx_expired(unsigned long long &t0, unsigned spins) {
#ifdef EQ_HOST_EMU
    (void)t0;
    return spins >= LSX_SPIN_LIMIT;
#else
    (void)spins;
    const unsigned long long now = lsx_gtime();
    if (t0 == 0) t0 = now;
    return now - t0 > LSX_TIMEOUT_NS;
#endif
}
#define LSX_TRACE(ev, q) do { if (p.trace && lane == 0 && k == p.K - 1 && (b < 3 || b == p.NB - 1) && (q) >= 200 && (q) < 328) p.trace[((size_t)(b < 3 ? b : 3) * 8 + (ev)) * 128 + (q) - 200] = lsx_gtime(); } while (0)
#define LSX_STAT(slot, v) do { if (p.stats && lane == 0) atomicAdd(p.stats + (slot), (unsigned long long)(v)); } while (0)

// ---- waiting primitives: every loop can be aborted ----
// ALL 32 lanes poll (the way the consumer warps of a TMA pipeline do).  A version where lane 0 polled
// and broadcast the result left the warp split 1 + 31 in the kernel's main loops -- ncu showed every
// step issued twice (16 threads per instruction on average) with each shuffle taking the
// WARPSYNC.COLLECTIVE slow path.  Polling with the whole warp has no lane-dependent branch at all.
// Dependency flags of other jobs (global memory, acquire).
__device__ __forceinline__ bool lsx_wait_flags(const unsigned *f1, unsigned n1, bool sys1, const unsigned *f2, unsigned n2,
                                               bool sys2, int *error, int lane) {
    bool ok = true;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    // spin with relaxed loads (an acquire load is followed by an L1 invalidate, CCTL.IVALL, which
    // the five loaders of an SM would otherwise issue every few hundred cycles), then take one
    // acquire load of the flag that was seen set: it reads from the release and synchronises.
    // (a flag written by the neighbouring GPU is read at system scope)
    while ((f1 && (sys1 ? ld_relaxed_sys_u32(f1) : ld_relaxed_u32(f1)) < n1) ||
           (f2 && (sys2 ? ld_relaxed_sys_u32(f2) : ld_relaxed_u32(f2)) < n2)) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0) {
            if (lsx_expired(t0, spins)) {
                if (lane == 0) *error = 1;
                ok = false;
                break;
            }
            if (ld_volatile_s32(error) != 0) {
                ok = false;
                break;
            }
        }
    }
    if (ok) {
        if (f1) (void)(sys1 ? ld_acquire_sys_u32(f1) : ld_acquire_u32(f1));
        if (f2) (void)(sys2 ? ld_acquire_sys_u32(f2) : ld_acquire_u32(f2));
    }
    return __all_sync(0xffffffffu, ok) != 0;
}
// An mbarrier of this CTA.
__device__ __forceinline__ bool lsx_wait_bar(uint32_t bar, uint32_t parity, int *error, int lane) {
    bool ok = true;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            if (lsx_expired(t0, spins)) {
                if (lane == 0) *error = 2;
                ok = false;
                break;
            }
            if (ld_volatile_s32(error) != 0) {
                ok = false;
                break;
            }
        }
    }
    return __all_sync(0xffffffffu, ok) != 0;
}

// Everything the three roles of a job share.
template <int ORIENT>
struct LsxJob {
    const LsxParams &p;
    const LsxProblem &pr;
    uint32_t sbase;
    int b, k, lane;
    int N, P, NC, NB, j0, M;
    bool last_band;
    bool first_local, last_local;   // my slab's first / last band has its neighbour on another GPU
    const uint8_t *cflags;

    __device__ __forceinline__ LsxJob(const LsxParams &p_, const LsxProblem &pr_, uint32_t sbase_, int b_, int k_,
                                      int lane_)
        : p(p_), pr(pr_), sbase(sbase_), b(b_), k(k_), lane(lane_) {
        N = p.N; P = p.P; NC = p.NC; NB = p.NB;
        j0 = 1 + 32 * b;
        M = (N + 31 + LSX_CW - 1) / LSX_CW;           // steps 0 .. N+30 in macro steps of LSX_CW
        last_band = (b == NB - 1);
        first_local = (b == p.b_lo && b > 0);
        last_local = (b == p.b_hi - 1 && b < NB - 1);
        cflags = p.chunk_flags + (ORIENT == EQ_ADJUST_COLUMN ? (size_t)NB * NC : 0) + (size_t)b * NC;
    }
    __device__ __forceinline__ uint32_t bar_full(int q) const { return sbase + LSX_BAR_OFF + (uint32_t)(q % LSX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t bar_done(int q) const { return sbase + LSX_BAR_OFF + (uint32_t)(LSX_SLOTS + q % LSX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t bar_free(int q) const { return sbase + LSX_BAR_OFF + (uint32_t)(2 * LSX_SLOTS + q % LSX_SLOTS) * 16u; }
    __device__ __forceinline__ uint32_t use_parity(int q) const { return (uint32_t)((q / LSX_SLOTS) & 1); }
    // How macro step m (steps CW*m .. CW*m+CW-1; columns CW*m-31 .. CW*m+CW-1) is executed:
    //   EDGE   some lane is on/next to a frame column: fully general loop (first and last few macro steps)
    //   CODED  interior columns, but the chunks it finalises hold fix-up codes of this orientation, or
    //          (Passive) the band owns a frame row: branch-light loop that reads one code byte per step
    //   FAST   interior columns, nothing to fix up: 3 LDS + 1 SHFL + 6 FP + 1 STS per step
    // Every job runs at the pace of the slowest job it depends on, so CODED must stay close to FAST.
    enum { MODE_FAST = 0, MODE_CODED = 1, MODE_EDGE = 2 };
    __device__ __forceinline__ int macro_mode(int m) const {
        if (LSX_CW * m - 31 < 2 || LSX_CW * m + LSX_CW - 1 > N - 2) return MODE_EDGE;
        if (ORIENT == EQ_PASSIVE) return MODE_FAST;   // the frame rows are blended in by the storer
        unsigned any = 0;
#pragma unroll
        for (int d = 0; d <= LSX_BACK; ++d) any |= cflags[m - d];     // codes are read in chunks m-BACK .. m
        return any ? MODE_CODED : MODE_FAST;
    }
    // the code tile of chunk q is read by macro steps q .. q+LSX_BACK
    __device__ __forceinline__ bool need_codes(int q) const {
        if (ORIENT == EQ_PASSIVE) return b == 0 || last_band;      // row 0 of the tile carries col_fluid
        bool need = false;
#pragma unroll
        for (int d = 0; d <= LSX_BACK; ++d) need = need || (macro_mode(q + d) != MODE_FAST);
        return need;
    }

    // ------------------------------------------------------------------ LOADER warp
    __device__ __forceinline__ bool run_loader() const {
        const float *__restrict__ x = pr.x;
        const float *__restrict__ x0 = pr.x0;
        const unsigned *flag_prev_iter = (k > 0) ? pr.progress + (size_t)(k - 1) * NB + min(b + 1, NB - 1) : nullptr;
        const unsigned *flag_band_above = (b > 0) ? pr.progress + (size_t)k * NB + (b - 1) : nullptr;
        const float *top_src = (b > 0) ? pr.raw + (size_t)b * P : x;  // row j0-1: raw stream or frame row 0
        constexpr int LPR = LSX_CW / 4;            // lanes per row (16 B each)
        constexpr int RPP = 32 / LPR;              // rows per pass
        const int sub = lane % LPR, rr = lane / LPR;
        for (int q = 0; q < NC; ++q) {
            // x0 is re-read from HBM every iteration and has no dependency: pull the chunk that will be
            // staged LSX_PF chunks from now into L2, one row per lane (in iteration 0 x is cold as well)
            if (q + LSX_PF < NC) {
                prefetch_l2(x0 + (size_t)(j0 + lane) * P + LSX_CW * (q + LSX_PF));
                if (k == 0) prefetch_l2(x + (size_t)(j0 + lane) * P + LSX_CW * (q + LSX_PF));
            }
            const long long t0 = p.stats ? lsx_clock() : 0;
            // the ring slot must have been written back (chunk q-SLOTS) ...
            if (q >= LSX_SLOTS && !lsx_wait_bar(bar_free(q), use_parity(q - LSX_SLOTS), p.error, lane)) return false;
            const long long t1 = p.stats ? lsx_clock() : 0;
            LSX_TRACE(0, q);
            // ... and the producers of this chunk must have published it
            if (!p.debug_nodeps &&
                !lsx_wait_flags(flag_prev_iter, (unsigned)min(q + 1 + p.slack, NC), last_local, flag_band_above,
                                (unsigned)min(q + 1 + p.slack, NC), first_local, p.error, lane))
                return false;
            if (p.stats) { const long long t2 = lsx_clock(); LSX_STAT(7, t1 - t0); LSX_STAT(8, t2 - t1); }
            const uint32_t slot = (uint32_t)(q % LSX_SLOTS) * (LSX_CW * 4u);   // byte offset of the chunk in a 512 B row
            const int col0 = LSX_CW * q;
            // x rows j0-1 .. j0+32 and x0 rows j0 .. j0+31
#pragma unroll
            for (int g = 0; g < (LSX_XROWS + RPP - 1) / RPP; ++g) {
                const int t = RPP * g + rr;
                if (t < LSX_XROWS) {
                    const float *src = (t == 0) ? top_src + col0 + 4 * sub
                                                : x + (size_t)(j0 - 1 + t) * P + col0 + 4 * sub;
                    cp_async_16s(sbase + LSX_XS_OFF + (uint32_t)t * 512u + slot + 16u * sub, src);
                }
            }
#pragma unroll
            for (int g = 0; g < 32 / RPP; ++g) {
                const int t = RPP * g + rr;
                cp_async_16s(sbase + LSX_X0_OFF + (uint32_t)t * 512u + slot + 16u * sub,
                             x0 + (size_t)(j0 + t) * P + col0 + 4 * sub);
            }
            // codes rows j0-1 .. j0+31, skipped when every macro step that touches this chunk takes the
            // fast loop.  Passive has no codes; a band that owns a frame row stages col_fluid (quirk Q6)
            // in row 0 of the code tile instead.
            if (need_codes(q)) {
                constexpr int CLPR = LSX_CW / 16;   // lanes per code row
                const uint32_t cslot = (uint32_t)(q % LSX_SLOTS) * LSX_CW;
                if (ORIENT == EQ_PASSIVE) {
                    if (lane < CLPR) cp_async_16s(sbase + LSX_CS_OFF + cslot + 16u * lane, p.col_fluid + col0 + 16 * lane);
                } else {
                    const int csub = lane % CLPR, crr = lane / CLPR;
#pragma unroll
                    for (int g = 0; g < (LSX_CROWS * CLPR + 31) / 32; ++g) {
                        const int t = (32 / CLPR) * g + crr;
                        if (t < LSX_CROWS)
                            cp_async_16s(sbase + LSX_CS_OFF + (uint32_t)t * 128u + cslot + 16u * csub,
                                         p.codes + (size_t)(j0 - 1 + t) * P + col0 + 16 * csub);
                    }
                }
            }
            LSX_TRACE(1, q);
            cp_async_mbar_arrive_noinc(bar_full(q));     // fires when this lane's copies have landed
        }
        return true;
    }

    // ------------------------------------------------------------------ STORER warp
    // Writes finished chunks back and frees their ring slots.  It never fences: the release
    // (which has to wait until the stores have drained, microseconds under load) is the
    // PUBLISHER's job, so the ring keeps moving while a release is in flight.
    __device__ __forceinline__ bool run_storer() const {
        float *__restrict__ x = pr.x;
        float *raw_out = (b + 1 < NB) ? (last_local ? pr.raw_down : pr.raw) + (size_t)(b + 1) * P : nullptr;
        const uint32_t raw_s = sbase + LSX_RAW_OFF;
        constexpr int LPR = LSX_CW / 4, RPP = 32 / LPR;
        const int sub = lane % LPR, rr = lane / LPR;
        // band rows (tile rows 1..32); the bottom frame row N-1 travels with the last band
        const int t_hi = last_band ? 33 : 32;
        const int t_lo = (ORIENT == EQ_PASSIVE && b == 0) ? 0 : 1;  // Passive rewrites frame row 0
        for (int q = 0; q < NC; ++q) {
            const long long t0 = p.stats ? lsx_clock() : 0;
            if (!lsx_wait_bar(bar_done(q), use_parity(q), p.error, lane)) return false;
            const long long t1 = p.stats ? lsx_clock() : 0;
            LSX_TRACE(4, q);
            const uint32_t slot = (uint32_t)(q % LSX_SLOTS) * (LSX_CW * 4u);
            const int col0 = LSX_CW * q;
#pragma unroll
            for (int g = 0; g < (LSX_XROWS + RPP - 1) / RPP; ++g) {
                const int t = RPP * g + rr;
                const int row = j0 - 1 + t;
                if (t >= t_lo && t <= t_hi && row <= N - 1) {
                    float4 v = lds_f32x4(sbase + LSX_XS_OFF + (uint32_t)t * 512u + slot + 16u * sub);
                    if (ORIENT == EQ_PASSIVE && (row == 0 || row == N - 1)) {
                        // Passive frame rows (fluid.rs:182-183): x[i,0] = x[i,1], x[i,N-1] = x[i,N-2] for the
                        // columns 1..N-2 that contain a NoWall cell (quirk Q6; col_fluid is staged in row 0
                        // of the code tile); interior rows are never changed by the Passive pass, so the
                        // neighbour row of the tile already holds the value to copy.
                        const int tn = (row == 0) ? t + 1 : t - 1;
                        const float4 nb = lds_f32x4(sbase + LSX_XS_OFF + (uint32_t)tn * 512u + slot + 16u * sub);
                        const uint32_t cf4 = lds_u32(sbase + LSX_CS_OFF + (uint32_t)(q % LSX_SLOTS) * LSX_CW + 4u * sub);
                        const int c0 = col0 + 4 * sub;
                        if ((cf4 & 0xffu) && c0 >= 1 && c0 <= N - 2) v.x = nb.x;
                        if ((cf4 & 0xff00u) && c0 + 1 >= 1 && c0 + 1 <= N - 2) v.y = nb.y;
                        if ((cf4 & 0xff0000u) && c0 + 2 >= 1 && c0 + 2 <= N - 2) v.z = nb.z;
                        if ((cf4 & 0xff000000u) && c0 + 3 >= 1 && c0 + 3 <= N - 2) v.w = nb.w;
                    }
                    *reinterpret_cast<float4 *>(x + (size_t)row * P + col0 + 4 * sub) = v;
                }
            }
            if (raw_out && lane < LSX_CW) raw_out[col0 + lane] = lds_f32(raw_s + (uint32_t)((col0 + lane) & 127) * 4u);
            if (first_local && lane < LPR) {
                // my first row is row j0+32 of the last band on rank-1: keep its copy there current
                const float4 v = lds_f32x4(sbase + LSX_XS_OFF + 512u + slot + 16u * lane);
                *reinterpret_cast<float4 *>(pr.x_up + (size_t)j0 * P + col0 + 4 * lane) = v;
            }
            __syncwarp();                             // every lane's smem reads and global stores are issued
            if (lane == 0) {
                mbar_arrive(bar_free(q));             // the ring slot may be refilled
                sts_release_cta_u32(sbase + LSX_MISC_OFF + 4u, (uint32_t)q + 1u);   // the publisher may release it
            }
            LSX_TRACE(5, q);
            if (p.stats) { const long long t2 = lsx_clock(); LSX_STAT(10, t1 - t0); LSX_STAT(11, t2 - t1); }
        }
        return true;
    }

    // ------------------------------------------------------------------ PUBLISHER warp
    // One lane: wait until the storer has issued the stores of the next chunk(s), then
    // st.release the progress counter.  The storer's stores happen-before the release through
    // the mbarrier (CTA scope) and the release is cumulative, the same pattern as
    // "__syncthreads(); if (tid == 0) { __threadfence(); flag = 1; }".
    __device__ __forceinline__ bool run_publisher() const {
        unsigned *my_flag = pr.progress + (size_t)k * NB + b;
        const uint32_t cnt = sbase + LSX_MISC_OFF + 4u;
        int q = 0, ok = 1;
        while (q < NC && ok) {
            if (lane == 0) {
                unsigned spins = 0;
                unsigned long long t0 = 0;
                int have;
                while ((have = (int)lds_acquire_cta_u32(cnt)) < min(q + p.pub_batch, NC)) {     // chunks whose stores are issued
                    __nanosleep(64);
                    if ((++spins & 1023u) == 0) {
                        if (lsx_expired(t0, spins)) { *p.error = 3; ok = 0; break; }
                        if (ld_volatile_s32(p.error) != 0) { ok = 0; break; }
                    }
                }
                if (ok) {
                    q = have;                                               // take along everything that is ready
                    const long long t2 = p.stats ? lsx_clock() : 0;
                    LSX_TRACE(6, q - 1);
                    st_release_u32(my_flag, (unsigned)q);
                    if (first_local) st_release_sys_u32(pr.prog_up + (size_t)k * NB + b, (unsigned)q);
                    if (last_local) st_release_sys_u32(pr.prog_down + (size_t)k * NB + b, (unsigned)q);
                    LSX_TRACE(7, q - 1);
                    if (p.stats) { const long long t3 = lsx_clock(); LSX_STAT(12, t3 - t2); LSX_STAT(9, 1); }
                }
            }
            q = __shfl_sync(0xffffffffu, q, 0);
            ok = __shfl_sync(0xffffffffu, ok, 0);
        }
        return ok != 0;
    }

    // ------------------------------------------------------------------ COMPUTE warp
    __device__ __forceinline__ bool run_compute() const {
        const int j = j0 + lane;
        const int tr = lane + 1;
        const bool in_row = (j <= N - 2);
        const float a = pr.a, c_recip = pr.c_recip;
        float *__restrict__ x = first_local ? pr.x_up : pr.x;   // only used for the cross-band DOWN patch of row j0-1
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
            // macro step m reads chunks m-LSX_BACK .. m+1
            const long long tw0 = p.stats ? lsx_clock() : 0;
            if (m + 1 < NC && !lsx_wait_bar(bar_full(m + 1), use_parity(m + 1), p.error, lane)) return false;
            const long long tw1 = p.stats ? lsx_clock() : 0;
            LSX_TRACE(2, m);

            const int mode = macro_mode(m);
            if (mode == MODE_FAST) {
                // ---- fast loop: every lane on an interior column, nothing to fix up.  Operands of
                // step t+1 are fetched before the arithmetic of step t (they are not written by
                // anyone before step t+3), so only SHFL -> 4 FP ops remain on the step-to-step chain.
                uint32_t o = ((uint32_t)(LSX_CW * m - lane) & 127u) << 2;  // byte offset of column c
                uint32_t om1 = (o - 4u) & 508u;                            // column c-1
                float right = lds_f32(xs_row + ((o + 4u) & 508u));
                float down = lds_f32(xs_row + 512u + o);
                float x0v = lds_f32(x0_row + o);
                float topv = (lane == 0) ? lds_f32(xs_top + o) : 0.f;
#pragma unroll
                for (int t = 0; t < LSX_CW; ++t) {
                    const uint32_t o1 = (o + 4u) & 508u, o2 = (o + 8u) & 508u;
                    float up = __shfl_up_sync(0xffffffffu, cur, 1);
                    if (lane == 0) up = topv;
                    const float right_n = lds_f32(xs_row + o2);
                    const float down_n = lds_f32(xs_row + 512u + o1);
                    const float x0_n = lds_f32(x0_row + o1);
                    if (lane == 0) topv = lds_f32(xs_top + o1);
                    const float newv = gs_update(x0v, right, cur, down, up, a, c_recip);
                    if (in_row) sts_f32(xs_row + om1, cur);                // column c-1 is final: F = R
                    if (lane == 31) sts_f32(raw_s + o, newv);
                    prev2 = cur;
                    prev_up = up;
                    cur = newv;
                    om1 = o;
                    o = o1;
                    right = right_n;
                    down = down_n;
                    x0v = x0_n;
                    __syncwarp();
                }
            } else {
                // ---- general loop, written with selects instead of branches: every shared-memory
                // address is valid for any column (the ring wraps), so operands are loaded
                // unconditionally and the column-range tests only pick results.  MODE_CODED skips
                // the range tests (every lane is on an interior column).
                const bool ranged = (mode == MODE_EDGE);
                const int s_end = min(LSX_CW * m + LSX_CW, S);
                for (int s = LSX_CW * m; s < s_end; ++s) {
                    const int c = s - lane;          // column this lane computes now (0 = left frame cell)
                    const int cf = c - 1;            // column finalised now
                    const uint32_t o = ((uint32_t)c & 127u) << 2;
                    const uint32_t om1 = (o - 4u) & 508u;
                    const float up = __shfl_up_sync(0xffffffffu, cur, 1);
                    const float right = lds_f32(xs_row + ((o + 4u) & 508u));
                    const float down = lds_f32(xs_row + 512u + o);
                    const float x0v = lds_f32(x0_row + o);
                    const float self = lds_f32(xs_row + o);
                    const float top = (lane == 0) ? lds_f32(xs_top + o) : up;
                    const bool gs_ok = in_row && (!ranged || (c >= 1 && c <= N - 2));
                    const bool pass_ok = (!ranged || (c >= 0 && c <= N - 1)) && (in_row || ORIENT == EQ_ADJUST_COLUMN);
                    const float g = gs_update(x0v, right, cur, down, top, a, c_recip);
                    const float newv = gs_ok ? g : (pass_ok ? self : cur);   // frame cells pass through
                    const bool fin = in_row && (!ranged || (cf >= 1 && cf <= N - 2));
                    float F = cur;                   // R_k(cf, j)
                    if (ORIENT == EQ_ADJUST_ROW) {
                        const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 3u;
                        F = (code == EQ_CODE_ROW_RIGHT) ? -newv : ((code == EQ_CODE_ROW_LEFT) ? -prev2 : cur);
                    } else if (ORIENT == EQ_ADJUST_COLUMN) {
                        const float dn = __shfl_down_sync(0xffffffffu, newv, 1);
                        const unsigned code = lds_u8(cs_row + (om1 >> 2)) & 12u;
                        const float below = (lane < 31) ? dn : lds_f32(xs_top + 33u * 512u + om1);
                        // a DOWN cell in lane 31 of a band that is not the last one is patched by the band below
                        const bool take_down = (code == EQ_CODE_COL_DOWN) && (lane < 31 || last_band);
                        F = (code == EQ_CODE_COL_UP) ? -prev_up : (take_down ? -below : cur);
                        if (gs_ok && lane == 0 && b > 0) {
                            // cell (c, j0-1) of the band above takes -R_k(c, j0) when its code says DOWN
                            const unsigned code0 = lds_u8(cs_top + (o >> 2)) & 12u;
                            if (code0 == EQ_CODE_COL_DOWN) x[(size_t)(j0 - 1) * P + c] = -newv;
                        }
                    }
                    if (fin) sts_f32(xs_row + om1, F);
                    if (ORIENT == EQ_PASSIVE && ranged && fin && row_has_fluid) {
                        // x[0,j] = x[1,j], x[N-1,j] = x[N-2,j] (fluid.rs:185-186, conditional per quirk Q6);
                        // the frame ROWS are blended in by the storer
                        if (cf == 1) sts_f32(xs_row, cur);
                        if (cf == N - 2) sts_f32(xs_row + (((uint32_t)(N - 1) & 127u) << 2), cur);
                    }
                    if (gs_ok && lane == 31) sts_f32(raw_s + o, newv);
                    prev2 = cur;
                    prev_up = top;
                    cur = newv;
                    __syncwarp();
                }
            }
            // every lane has finalised the columns below CW*(m+1)-32: hand that chunk to the storer
            if (m >= LSX_BACK && m - LSX_BACK < NC && lane == 0) mbar_arrive(bar_done(m - LSX_BACK));
            LSX_TRACE(3, m);
            if (p.stats) { const long long tw2 = lsx_clock(); LSX_STAT(0, tw1 - tw0); LSX_STAT(1 + mode, tw2 - tw1); LSX_STAT(4 + mode, 1); }
        }
        if (lane == 0)
            for (int q = max(M - LSX_BACK, 0); q < NC; ++q) mbar_arrive(bar_done(q));
        return true;
    }
};

// Four warps per CTA (compute / loader / storer / publisher); persistent CTAs pull jobs from the ticket counter.
__global__ void __launch_bounds__(LSX_THREADS) k_linsolve_exact(const LsxParams p) {
    if (p.run_if && *p.run_if == 0u) return;
    EQ_DYN_SMEM(lsx_smem_raw);
    const uint32_t sbase = smem_u32(lsx_smem_raw);
    const int total = p.njobs * p.nprob;
    // broadcast the warp index so the compiler knows the role branches below are warp-uniform
    // (otherwise every __shfl/__syncwarp in the roles becomes an out-of-line WARPSYNC.COLLECTIVE)
    const int lane = (int)threadIdx.x & 31;
    if (threadIdx.x == 0) sts_u32(sbase + LSX_MISC_OFF + 8u, p.rotate_roles ? eq_cta_slot_rotation() : 0u);
    __syncthreads();
    // role 0 compute, 1 loader, 2 storer, 3 publisher (uniform per warp: taken through a shuffle so that
    // the compiler keeps the dispatch branch-uniform)
    const int warp = __shfl_sync(0xffffffffu, (((int)threadIdx.x >> 5) - (int)lds_u32(sbase + LSX_MISC_OFF + 8u)) & 3, 0);
    for (;;) {
        __syncthreads();                               // the previous job is finished in all three roles
        if (threadIdx.x == 0) {
            const unsigned t = (ld_volatile_s32(p.error) != 0) ? 0xffffffffu : atomicAdd(p.ticket, 1u);
            sts_u32(sbase + LSX_MISC_OFF, t);
            sts_u32(sbase + LSX_MISC_OFF + 4u, 0u);
            // (re-initialised per job: only sound while a job completes an even number of phases per slot, see single_shot)
            for (int i = 0; i < LSX_SLOTS; ++i) {
                mbar_init(sbase + LSX_BAR_OFF + (uint32_t)i * 16u, 32u);                       // full: 32 loader lanes
                mbar_init(sbase + LSX_BAR_OFF + (uint32_t)(LSX_SLOTS + i) * 16u, 1u);          // done: compute lane 0
                mbar_init(sbase + LSX_BAR_OFF + (uint32_t)(2 * LSX_SLOTS + i) * 16u, 1u);      // free: storer lane 0
            }
        }
        __syncthreads();
        const unsigned t = lds_u32(sbase + LSX_MISC_OFF);
        if (t >= (unsigned)total) break;
        const int pi = (int)(t % (unsigned)p.nprob);
        const uint32_t jb = p.jobs[t / (unsigned)p.nprob];
        const int k = (int)(jb >> 16), b = (int)(jb & 0xffffu);
        const LsxProblem &pr = p.prob[pi];
        if (p.jobtimes && threadIdx.x == 0) p.jobtimes[2 * ((size_t)(k * p.NB + b) * p.nprob + pi)] = lsx_gtime();
#define LSX_DISPATCH(O)                                           \
    {                                                             \
        const LsxJob<O> job(p, pr, sbase, b, k, lane);            \
        if (warp == 0) job.run_compute();                         \
        else if (warp == 1) job.run_loader();                     \
        else if (warp == 2) job.run_storer();                     \
        else job.run_publisher();                                 \
    }
        if (pr.orient == EQ_ADJUST_ROW) LSX_DISPATCH(EQ_ADJUST_ROW)
        else if (pr.orient == EQ_ADJUST_COLUMN) LSX_DISPATCH(EQ_ADJUST_COLUMN)
        else LSX_DISPATCH(EQ_PASSIVE)
#undef LSX_DISPATCH
        if (p.jobtimes && threadIdx.x == 0) p.jobtimes[2 * ((size_t)(k * p.NB + b) * p.nprob + pi) + 1] = lsx_gtime();
        if (p.single_shot) break;
        // a role that gave up has set *p.error (or seen it set); the next pass of the loop makes
        // thread 0 read it and every thread leaves together
    }
}
