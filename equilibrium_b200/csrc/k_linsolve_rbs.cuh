// k_linsolve_rbs.cuh -- red-black lin_solve (EQ_MODE_RED_BLACK) as a SLIDING WINDOW over the rows.
//
// k_rb_reg keeps a 256 x 64 tile in registers and runs 4 iterations on it: 12 halo cells on every side, so a tile
// writes 232 x 40 of the 256 x 64 cells it updates (1.77 x the useful work), x0 lives in shared memory (one LDS per
// update) and the unrolled body is 21 KB (ncu: 26 % stall_no_inst, 0.22 of the HBM peak).  Here a WARP owns a strip of
// 128 columns (4 per lane) and streams down the rows of a segment:
//
//   tick tau:   R_0 at row tau,  B_0 at tau-1,  F_0 at tau-2,  R_1 at tau-3,  B_1 at tau-4,  F_1 at tau-5
//               (R / B = red / black half-sweep, F = set_boundaries of iteration t; each stage reads the rows above and
//               below its own after the stage before it has passed them), row tau-5 leaves, row tau+2 arrives
//
// so only the strip's 8 halo columns per side and 6 rows at the ends of a segment are recomputed (1.16 x), x AND x0
// stay in registers (a rotating window of 8 rows, static indices after unrolling 8 ticks), a half-sweep of a row costs
// one shuffle (the lane-crossing neighbour of the row's first or last active column), warps never synchronise with each
// other and the kernel has no shared memory.  Two iterations per pass: the pass is a float4 stream of x, x0 in and x
// out -- HBM-bound by construction.
//
// Bit-identical to k_rb_reg and to the oracle's red-black restatement (same expression tree, same colour order, same
// set_boundaries); results go to the ping-pong partner like k_rb_reg's.  Mirror codes come from codes[] (AdjustRow bits
// 0-1, AdjustColumn bits 2-3, the Passive frame copies bits 4-6).
#pragma once
#include "eq_common.cuh"
#include <type_traits>

#define RS_T 2                       // iterations per pass
#define RS_HALO 8                    // strip columns on each side that are not written (>= 3 RS_T, a multiple of 4)
#define RS_SW (128 - 2 * RS_HALO)    // 112 output columns per strip
#define RS_VH (3 * RS_T)             // rows recomputed at both ends of a segment
#define RS_WARPS 4
#define RS_THREADS (32 * RS_WARPS)
#define RS_SEG 256                   // most rows per segment (task = strip x segment); fewer on small grids, see the host code
#ifndef RS_PF
#define RS_PF 12                     // rows ahead of the stream that are pulled into L2
#endif

template <int ORIENT>
__device__ __forceinline__ unsigned rs_decode(unsigned byte) {
    if (ORIENT == EQ_ADJUST_ROW) return byte & 3u;                                                     // 1 = L, 2 = R
    if (ORIENT == EQ_ADJUST_COLUMN) { const unsigned c = (byte >> 2) & 3u; return c ? c + 2u : 0u; }   // 3 = U, 4 = D
    return (byte >> EQ_CODE_PASSIVE_SHIFT) & 7u;
}

// GUARD = false: every row the task touches (ys - VH - 9 .. ye + VH + 8) and every column of the strip is interior:
// no range tests on loads, updates or stores.
template <int ORIENT, int ITERS, bool GUARD>
__device__ __forceinline__ void rs_task(const float *__restrict__ xin, float *__restrict__ xout, const float *__restrict__ x0,
                                        const uint8_t *__restrict__ codes, bool need_fix, float a, float c_recip,
                                        int X0c, int ys, int ye, const EqLayout &L, int lane) {
    const int N = L.N, P = L.P;
    const int c0 = X0c + 4 * lane;                                  // my four columns c0 .. c0+3
    const bool ld_ok = (c0 >= 0 && c0 < P);                         // (c0 is a multiple of 4 and P of 32: all four or none)
    bool colok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) colok[e] = (c0 + e >= 1 && c0 + e <= N - 2);
    const bool st_lane = (lane >= RS_HALO / 4 && lane < 32 - RS_HALO / 4) && c0 < N;
    const bool st_full = st_lane && (c0 + 3 < N);
    constexpr unsigned SGN = (ORIENT != EQ_PASSIVE) ? 0x80000000u : 0u;

    float X[8][4], Z[8][4];                                         // rows y of x / x0 live in slot y & 7
    unsigned CW[8];                                                  // their mirror codes, one byte per cell
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        CW[s] = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) X[s][e] = Z[s][e] = 0.f;
    }
    auto load_row = [&](int y, float (&xr)[4], float (&zr)[4], unsigned &cw) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), zv = xv;
        cw = 0u;
        if (!GUARD || (ld_ok && y >= 0 && y < N)) {
            xv = *reinterpret_cast<const float4 *>(xin + (size_t)y * P + c0);
            zv = *reinterpret_cast<const float4 *>(x0 + (size_t)y * P + c0);
            if (GUARD && need_fix) cw = *reinterpret_cast<const unsigned *>(codes + (size_t)y * P + c0);
        }
        xr[0] = xv.x; xr[1] = xv.y; xr[2] = xv.z; xr[3] = xv.w;
        zr[0] = zv.x; zr[1] = zv.y; zr[2] = zv.z; zr[3] = zv.w;
    };
    // one half-sweep of row y (colour C): the active cells are the e with (e + y + C) even (c0 is even)
    // (slot numbers and parities are passed in as functions of the unrolled tick index: static after inlining)
    auto half_sweep = [&](int y, int s /* y & 7 */, int par /* (y + C) & 1 */) {
        const int su = (s + 7) & 7, sd = (s + 1) & 7;
        const bool rowok = !GUARD || (y >= 1 && y <= N - 2);
        if (par == 0) {                                             // e = 0, 2
            const float l0 = __shfl_up_sync(0xffffffffu, X[s][3], 1);
            const float n0 = gs_update(Z[s][0], X[s][1], l0, X[sd][0], X[su][0], a, c_recip);
            const float n2 = gs_update(Z[s][2], X[s][3], X[s][1], X[sd][2], X[su][2], a, c_recip);
            X[s][0] = (!GUARD || (rowok && colok[0])) ? n0 : X[s][0];
            X[s][2] = (!GUARD || (rowok && colok[2])) ? n2 : X[s][2];
        } else {                                                    // e = 1, 3
            const float r3 = __shfl_down_sync(0xffffffffu, X[s][0], 1);
            const float n1 = gs_update(Z[s][1], X[s][2], X[s][0], X[sd][1], X[su][1], a, c_recip);
            const float n3 = gs_update(Z[s][3], r3, X[s][2], X[sd][3], X[su][3], a, c_recip);
            X[s][1] = (!GUARD || (rowok && colok[1])) ? n1 : X[s][1];
            X[s][3] = (!GUARD || (rowok && colok[3])) ? n3 : X[s][3];
        }
    };
    // set_boundaries on row y: every coded cell takes (minus) its neighbour; sources are never destinations
    auto fix_row = [&](int s /* y & 7 */) {
        const int su = (s + 7) & 7, sd = (s + 1) & 7;
        const float lft = __shfl_up_sync(0xffffffffu, X[s][3], 1), rgt = __shfl_down_sync(0xffffffffu, X[s][0], 1);
        const unsigned cw = CW[s];
        float nv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned d = rs_decode<ORIENT>((cw >> (8 * e)) & 255u);
            float src = X[s][e];
            src = (d == WF_C_L) ? (e == 0 ? lft : X[s][e > 0 ? e - 1 : 0]) : src;
            src = (d == WF_C_R) ? (e == 3 ? rgt : X[s][e < 3 ? e + 1 : 3]) : src;
            src = (d == WF_C_U) ? X[su][e] : src;
            src = (d == WF_C_D) ? X[sd][e] : src;
            nv[e] = (d != WF_C_NONE) ? __uint_as_float(__float_as_uint(src) ^ SGN) : src;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) X[s][e] = nv[e];
    };

    // ticks: R_0 first touches row ys - VH (rows above feed garbage that dies inside the 3-rows-per-iteration cone);
    // the loop starts at a multiple of 8 so that slot numbers are static inside the unrolled body
    const int t_first = ys - RS_VH, t_last = ye - 1 + 3 * ITERS - 1;
    int tb = (t_first >= 0) ? (t_first & ~7) : -((-t_first + 7) & ~7);
    // rows tb-1, tb, tb+1 must be in the window when tick tb runs; row tb+2 arrives during it
    {
        float pz[4];
        load_row(tb - 1, X[7], Z[7], CW[7]);                         // tb is a multiple of 8
        load_row(tb, X[0], Z[0], CW[0]);
        load_row(tb + 1, X[1], Z[1], CW[1]);
        (void)pz;
    }
    for (; tb <= t_last; tb += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int tau = tb + k;                                   // tau & 7 == k, tau & 1 == k & 1
            float px[4], pz[4];
            unsigned pc;
            if (!GUARD || (ld_ok && tau + RS_PF < N)) {               // a warp has one row in flight in registers: hide DRAM behind L2
                prefetch_l2(xin + (size_t)(tau + RS_PF) * P + c0);
                prefetch_l2(x0 + (size_t)(tau + RS_PF) * P + c0);
            }
            load_row(tau + 2, px, pz, pc);                            // lands in slot (tau + 2) & 7 = slot of row tau - 6 (dead after this tick)
            half_sweep(tau, k & 7, (k + 0) & 1);                      // R_0: colour 0 cells of row tau
            half_sweep(tau - 1, (k + 7) & 7, (k + 7 + 1) & 1);        // B_0: colour 1 cells of row tau - 1
            if (GUARD && need_fix) fix_row((k + 6) & 7);              // F_0: row tau - 2 (tasks with mirror codes take the guarded path)
            if (ITERS > 1) {
                half_sweep(tau - 3, (k + 5) & 7, (k + 5 + 0) & 1);    // R_1: row tau - 3
                half_sweep(tau - 4, (k + 4) & 7, (k + 4 + 1) & 1);    // B_1: row tau - 4
                if (GUARD && need_fix) fix_row((k + 3) & 7);          // F_1: row tau - 5
            }
            const int yo = tau - 3 * ITERS + 1;                       // this row is final now
            if (yo >= ys && yo < ye) {
                constexpr int so_off = (8 * 4 - 3 * ITERS + 1);       // (k - 3 ITERS + 1) & 7
                const int so = (k + so_off) & 7;
                if (!GUARD ? st_lane : st_full) {
                    *reinterpret_cast<float4 *>(xout + (size_t)yo * P + c0) = make_float4(X[so][0], X[so][1], X[so][2], X[so][3]);
                } else if (st_lane) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c0 + e < N) xout[(size_t)yo * P + c0 + e] = X[so][e];
                }
            }
            const int sn = (k + 2) & 7;
#pragma unroll
            for (int e = 0; e < 4; ++e) { X[sn][e] = px[e]; Z[sn][e] = pz[e]; }
            CW[sn] = pc;
        }
    }
}

// grid.x * RS_WARPS >= strips * segments; one task per warp
#ifndef RS_CTAS_PER_SM
#define RS_CTAS_PER_SM 4
#endif
__global__ void __launch_bounds__(RS_THREADS, RS_CTAS_PER_SM) k_rb_slide(const float *__restrict__ xin, float *__restrict__ xout,
                                                         const float *__restrict__ x0, const uint8_t *__restrict__ codes,
                                                         const uint8_t *__restrict__ chunk_flags, float a, float c_recip, int orient,
                                                         int iters, int row_lo, int row_hi, int nstrips, int nsegs, int seg_rows,
                                                         const unsigned *__restrict__ run_if, EqLayout L) {
    if (run_if && *run_if == 0u) return;                                   // the a == 0 shortcut was taken (k_a0_check)
    const int lane = (int)threadIdx.x & 31;
    const int task = (int)blockIdx.x * RS_WARPS + ((int)threadIdx.x >> 5);
    if (task >= nstrips * nsegs) return;
    const int sx = task % nstrips, sy = task / nstrips;
    const int N = L.N;
    const int X0c = sx * RS_SW - RS_HALO;
    const int ys = row_lo + sy * seg_rows, ye = min(ys + seg_rows, row_hi);
    // does set_boundaries have anything to do in this task?  Passive: only along the frame.  AdjustRow / AdjustColumn:
    // the (band, chunk) summaries of k_build_codes (band = (row-1)/32, chunk = column / EQ_LSX_CW); a superset is fine
    bool need_fix;
    if (orient == EQ_PASSIVE) {
        need_fix = (X0c <= 0) || (X0c + 128 >= N) || (ys - RS_VH - 1 <= 0) || (ye + RS_VH + 1 >= N - 1);
    } else {
        const int NB = (N - 2 + 31) / 32, NC = (N + EQ_LSX_CW - 1) / EQ_LSX_CW;
        const int b_lo = max((max(ys - RS_VH - 1, 1) - 1) / 32, 0), b_hi = min((min(ye + RS_VH, N - 2) - 1) / 32 + 1, NB - 1);
        const int q_lo = max(X0c, 0) / EQ_LSX_CW, q_hi = min(X0c + 127, N - 1) / EQ_LSX_CW;
        const uint8_t *fl = chunk_flags + (orient == EQ_ADJUST_COLUMN ? (size_t)NB * NC : 0);
        const int nq = q_hi - q_lo + 1, total = max(b_hi - b_lo + 1, 0) * max(nq, 0);
        int any = 0;
        for (int t = lane; t < total; t += 32) any |= fl[(size_t)(b_lo + t / nq) * NC + q_lo + t % nq];
        need_fix = __any_sync(0xffffffffu, any != 0) != 0;
    }
    // interior task: no range tests (the window touches rows ys - VH - 9 .. ye + VH + 8 at most)
    const bool inside = X0c >= 1 && X0c + 127 <= N - 2 && ys - RS_VH - 9 >= 1 && ye + RS_VH + 8 + RS_PF <= N - 2;
#define RS_GO(O, I, G) rs_task<O, I, G>(xin, xout, x0, codes, need_fix, a, c_recip, X0c, ys, ye, L, lane)
    if (iters >= 2 && inside && !need_fix) {
        // (the mirror codes of a task without any are never looked at, so one instantiation serves every orientation)
        RS_GO(EQ_PASSIVE, 2, false);
    } else if (iters >= 2) {
        if (orient == EQ_ADJUST_ROW) RS_GO(EQ_ADJUST_ROW, 2, true);
        else if (orient == EQ_ADJUST_COLUMN) RS_GO(EQ_ADJUST_COLUMN, 2, true);
        else RS_GO(EQ_PASSIVE, 2, true);
    } else {
        if (orient == EQ_ADJUST_ROW) RS_GO(EQ_ADJUST_ROW, 1, true);
        else if (orient == EQ_ADJUST_COLUMN) RS_GO(EQ_ADJUST_COLUMN, 1, true);
        else RS_GO(EQ_PASSIVE, 1, true);
    }
#undef RS_GO
}

// ---------------------------------------------------------------------------------------------------------------------
// k_rb_stream: a sliding window with FOUR iterations per pass, the rows brought in by the TMA engine.
//
//   tick tau:   R_i at row tau - 3i,  B_i at tau - 3i - 1,  F_i at tau - 3i - 2   (i = 0..3),
//               row tau - 11 leaves, row tau + 1 is taken out of the x ring, row tau + 5 is requested
//
// * One warp per CTA and one task per warp: everything that describes the task is warp-uniform for the compiler (uniform
//   registers, uniform branches), and the four mbarriers are initialised once.
// * Every iteration ("stage") keeps its own ring of 4 rows in registers (slot = row & 3): the rows R_i, B_i and F_i work
//   on plus the row F_i finished one tick ago.  R_i takes the old values and the lower neighbour of its row from the
//   ring of stage i - 1 (stage 0: from the two raw rows taken last) and writes the row into its own ring -- two
//   new values and two copies.  The rotation period is 4 ticks, so the unrolled loop body is 4 ticks (~10 KB).  (One
//   16-slot window for all stages needs no copies but 16 unrolled ticks: 54 KB of code, and 16 independent warps per
//   SM at 16 different places of it -- ncu: 41 % of the stall samples were "no instruction".)
// * x0 stays in a per-warp ring of 16 rows in shared memory: a row is read once per iteration (R_i loads it, B_i of the
//   next tick reuses two registers).
// * Lane 0 requests row tau + 5 of x, x0 (and the mirror codes, when the task has any) with bulk copies
//   (cp.async.bulk, SASS UBLKCP) that complete on one of four mbarriers; every barrier completes one phase per row and
//   four rows are always in flight per warp -- no prefetch instructions, no address arithmetic in the other lanes.
// * A pass moves 12 bytes per cell for four iterations: 3 bytes per cell-iteration.
//
// Same expression tree, colour order and set_boundaries as k_rb_slide / k_rb_reg / the oracle: bit-identical results.
#define RQ_T 4
#define RQ_HALO 12                    // strip columns on each side that are not written (= 3 RQ_T, a multiple of 4)
#define RQ_SW (128 - 2 * RQ_HALO)     // 104 output columns per strip
#define RQ_VH (3 * RQ_T)              // rows recomputed at both ends of a segment
#define RQ_THREADS 32
#define RQ_CROW 144                   // bytes per row of the code ring: the strip's 128 codes widened to 16-byte boundaries
#define RQ_Z_OFF 0                    // x0 ring: 16 rows x 512 B
#define RQ_X_OFF 8192                 // x ring: 4 rows x 512 B
#define RQ_C_OFF 10240                // code ring: 16 rows x RQ_CROW
#define RQ_B_OFF (10240 + 16 * RQ_CROW)   // 4 mbarriers (16-byte slots)
#define RQ_WARP_BYTES (RQ_B_OFF + 64)     // 12608
#define RQ_SMEM_BYTES RQ_WARP_BYTES
#define RQ_BLOCK 16                   // ticks between two looks at the (band, chunk) flags
#ifndef RQ_CTAS_PER_SM
#define RQ_CTAS_PER_SM 16
#endif

// The ticks come in three variants and every RQ_BLOCK ticks pick their own:
//   <GUARD, FIX> = <false, false>  no range tests, no set_boundaries: everything away from the walls and the obstacles
//                  <false, true>   mirror codes in the rows the F stages touch ((band, chunk) flags of k_build_codes)
//                  <true, true>    range tests: the strips along the left / right wall, the ticks near rows 0 and N - 1
struct RqTask {
    // warp-uniform
    const float *__restrict__ xin;
    const float *__restrict__ x0;
    const uint8_t *__restrict__ codes;
    float *__restrict__ xout;
    float *push_up, *push_dn;                                        // the neighbours' copies of xout (row slabs), or nullptr
    int push_lo, push_hi;                                            // rows < push_lo go up as well, rows >= push_hi down
    unsigned char *wsm;
    float a, c_recip;
    int N, P, X0c, ys, ye, t_last, cb, qb;
    bool with_codes;
    unsigned ph;
    uint32_t bar0, fbytes, cbytes, xdst, zdst, cdst;
    const float *rqx, *rqz;                                          // the next row to request (tau + 5), first copied column
    const uint8_t *rqc;
    // per lane
    int lane, c0;
    unsigned ooff;                                                   // element offset of (row tau - 11, column c0) in xout
    const unsigned char *zring, *xring, *cring;
    float XL[2][4];                                                  // raw rows tau, tau + 1 (slot = row & 1)
    float S[RQ_T][4][4];                                             // stage i: rows tau - 3i .. tau - 3i - 3 (slot = row & 3)
    float Zc[RQ_T][2];                                               // x0 of the two cells that B_i sweeps in the next tick

    __device__ __forceinline__ void init(const EqLayout &L) {
        N = L.N;
        P = L.P;
        c0 = X0c + 4 * lane;
        ph = 0u;
        zring = wsm + RQ_Z_OFF + 16 * lane;
        xring = wsm + RQ_X_OFF + 16 * lane;
        const int abase = (X0c >= 0) ? (X0c & ~15) : -((-X0c + 15) & ~15);       // the code ring starts at this column
        cring = wsm + RQ_C_OFF + (c0 - abase);
        bar0 = smem_u32(wsm + RQ_B_OFF);
        // what lane 0 copies per row: the strip's columns clipped to the padded row
        cb = max(X0c, 0);
        const int ce = min(X0c + 128, P);
        qb = max(abase, 0);
        const int qe = min((X0c + 128 + 15) & ~15, P);
        fbytes = (uint32_t)(ce - cb) * 4u;
        cbytes = (uint32_t)(qe - qb);
        zdst = smem_u32(wsm + RQ_Z_OFF) + (uint32_t)(cb - X0c) * 4u;
        xdst = smem_u32(wsm + RQ_X_OFF) + (uint32_t)(cb - X0c) * 4u;
        cdst = smem_u32(wsm + RQ_C_OFF) + (uint32_t)(qb - abase);
        t_last = ye - 1 + 3 * RQ_T - 1;
#pragma unroll
        for (int e = 0; e < 4; ++e) XL[0][e] = XL[1][e] = 0.f;
#pragma unroll
        for (int i = 0; i < RQ_T; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) S[i][q][e] = 0.f;
            Zc[i][0] = Zc[i][1] = 0.f;
        }
    }
    // one lane: row y (rqx, rqz, rqc point at it) -> x ring slot y & 3 (= s4), x0 / code ring slot y & 15
    template <bool GUARD>
    __device__ __forceinline__ void request(int y, int s4) const {
        const uint32_t bar = bar0 + 16u * (uint32_t)s4;
        if (!GUARD || (y >= 0 && y < N)) {
            const uint32_t s16 = (uint32_t)y & 15u;
            bulk_g2s(xdst + 512u * (uint32_t)s4, rqx, fbytes, bar);
            bulk_g2s(zdst + 512u * s16, rqz, fbytes, bar);
            if (with_codes) {
                bulk_g2s(cdst + (uint32_t)RQ_CROW * s16, rqc, cbytes, bar);
                mbar_arrive_expect_tx(bar, 2u * fbytes + cbytes);
            } else {
                mbar_arrive_expect_tx(bar, 2u * fbytes);
            }
        } else {
            mbar_arrive(bar);                                        // a row that does not exist: nothing lands, the phase still ends
        }
    }
    __device__ __forceinline__ void await(int s4, unsigned par) const {
        unsigned spins = 0u;
        while (!mbar_try_wait(bar0 + 16u * (uint32_t)s4, par)) {
#ifndef EQ_HOST_EMU
            if (++spins > (1u << 26)) __trap();                       // a copy that never lands: fail the launch, do not hang the GPU
#endif
        }
        (void)spins;
    }
    __device__ __forceinline__ void take_x(int sl, int s4) {
        const float4 v = *reinterpret_cast<const float4 *>(xring + 512 * s4);
        XL[sl][0] = v.x; XL[sl][1] = v.y; XL[sl][2] = v.z; XL[sl][3] = v.w;
    }
    template <bool GUARD>
    __device__ __forceinline__ bool cell_ok(int y, int e) const {
        return !GUARD || (y >= 1 && y <= N - 2 && c0 + e >= 1 && c0 + e <= N - 2);
    }
    // R_i: the cells e = par, par + 2 of row y (c0 is even) from their old row `o`, the rows above and below; the whole
    // row goes to `d`
    template <bool GUARD>
    __device__ __forceinline__ void sweep_into(int y, int par, float za, float zb, const float (&o)[4], const float (&up)[4],
                                               const float (&dn)[4], float (&d)[4]) {
        if (par == 0) {
            const float l0 = __shfl_up_sync(0xffffffffu, o[3], 1);
            const float n0 = gs_update(za, o[1], l0, dn[0], up[0], a, c_recip);
            const float n2 = gs_update(zb, o[3], o[1], dn[2], up[2], a, c_recip);
            d[0] = cell_ok<GUARD>(y, 0) ? n0 : o[0];
            d[1] = o[1];
            d[2] = cell_ok<GUARD>(y, 2) ? n2 : o[2];
            d[3] = o[3];
        } else {
            const float r3 = __shfl_down_sync(0xffffffffu, o[0], 1);
            const float n1 = gs_update(za, o[2], o[0], dn[1], up[1], a, c_recip);
            const float n3 = gs_update(zb, r3, o[2], dn[3], up[3], a, c_recip);
            d[0] = o[0];
            d[1] = cell_ok<GUARD>(y, 1) ? n1 : o[1];
            d[2] = o[2];
            d[3] = cell_ok<GUARD>(y, 3) ? n3 : o[3];
        }
    }
    // B_i: in place
    template <bool GUARD>
    __device__ __forceinline__ void sweep(int y, int par, float za, float zb, float (&r)[4], const float (&up)[4], const float (&dn)[4]) {
        if (par == 0) {
            const float l0 = __shfl_up_sync(0xffffffffu, r[3], 1);
            const float n0 = gs_update(za, r[1], l0, dn[0], up[0], a, c_recip);
            const float n2 = gs_update(zb, r[3], r[1], dn[2], up[2], a, c_recip);
            r[0] = cell_ok<GUARD>(y, 0) ? n0 : r[0];
            r[2] = cell_ok<GUARD>(y, 2) ? n2 : r[2];
        } else {
            const float r3 = __shfl_down_sync(0xffffffffu, r[0], 1);
            const float n1 = gs_update(za, r[2], r[0], dn[1], up[1], a, c_recip);
            const float n3 = gs_update(zb, r3, r[2], dn[3], up[3], a, c_recip);
            r[1] = cell_ok<GUARD>(y, 1) ? n1 : r[1];
            r[3] = cell_ok<GUARD>(y, 3) ? n3 : r[3];
        }
    }
    // F_i: set_boundaries on row y: every coded cell takes (minus) its neighbour; sources are never destinations.
    // AdjustRow only knows LEFT / RIGHT, AdjustColumn only UP / DOWN (no shuffles), Passive (the frame) all four.
    template <bool GUARD, int ORIENT>
    __device__ __forceinline__ void fix_row(int y, float (&r)[4], const float (&up)[4], const float (&dn)[4]) {
        unsigned cw = *reinterpret_cast<const unsigned *>(cring + RQ_CROW * (y & 15));
        if (GUARD && !(c0 >= 0 && c0 < P && y >= 0 && y < N)) cw = 0u;
        constexpr unsigned VM = ORIENT == EQ_ADJUST_ROW ? 0x03030303u : (ORIENT == EQ_ADJUST_COLUMN ? 0x0c0c0c0cu : 0x70707070u);
        constexpr unsigned SGN = (ORIENT != EQ_PASSIVE) ? 0x80000000u : 0u;
        // most rows of a strip carry no code at all: one vote instead of the selects
        if (!__any_sync(0xffffffffu, (cw & VM) != 0u)) return;
        // branch-free: a cell's code bits become all-ones / zero masks, the cell takes the OR of its masked candidates
        // (ptxas compiled a chain of ?: into divergent branches with a reconvergence barrier per cell)
        unsigned lft = 0u, rgt = 0u;
        if (ORIENT != EQ_ADJUST_COLUMN) {
            lft = __float_as_uint(__shfl_up_sync(0xffffffffu, r[3], 1));
            rgt = __float_as_uint(__shfl_down_sync(0xffffffffu, r[0], 1));
        }
        const unsigned me[4] = {__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3])};
        unsigned nv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned b = cw >> (8 * e);
            unsigned mL = 0u, mR = 0u, mU = 0u, mD = 0u;
            if (ORIENT == EQ_ADJUST_ROW) {                            // bits 0-1: 1 = LEFT, 2 = RIGHT
                mL = 0u - (b & 1u);
                mR = 0u - ((b >> 1) & 1u);
            } else if (ORIENT == EQ_ADJUST_COLUMN) {                  // bits 2-3: 1 = UP, 2 = DOWN
                mU = 0u - ((b >> 2) & 1u);
                mD = 0u - ((b >> 3) & 1u);
            } else {                                                  // bits 4-6: WF_C_L .. WF_C_D
                const unsigned d = (b >> EQ_CODE_PASSIVE_SHIFT) & 7u;
                mL = 0u - (unsigned)(d == WF_C_L);
                mR = 0u - (unsigned)(d == WF_C_R);
                mU = 0u - (unsigned)(d == WF_C_U);
                mD = 0u - (unsigned)(d == WF_C_D);
            }
            const unsigned any = mL | mR | mU | mD;
            const unsigned vl = (e == 0) ? lft : me[e > 0 ? e - 1 : 0], vr = (e == 3) ? rgt : me[e < 3 ? e + 1 : 3];
            unsigned v = me[e] & ~any;
            if (ORIENT != EQ_ADJUST_COLUMN) v |= (vl & mL) | (vr & mR);
            if (ORIENT != EQ_ADJUST_ROW) v |= (__float_as_uint(up[e]) & mU) | (__float_as_uint(dn[e]) & mD);
            nv[e] = v ^ (SGN & any);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = __uint_as_float(nv[e]);
    }
    __device__ __forceinline__ void next_row() {
        ooff += (unsigned)P;
        rqx += P;
        rqz += P;
        rqc += P;
    }
    __device__ __forceinline__ void prologue(int tb) {
        const ptrdiff_t ro = (ptrdiff_t)tb * P;
        rqx = xin + (ro + cb);
        rqz = x0 + (ro + cb);
        rqc = codes + (ro + qb);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (elect_one()) request<true>(tb + q, q);
            next_row();
        }
        await(0, 0u);
        take_x(0, 0);
        __syncwarp();
        if (elect_one()) request<true>(tb + 4, 0);
        next_row();
        // (mod 2^32: the offset is only used once row tau - 11 >= ys >= 0; the host keeps N * P below 2^32)
        ooff = (unsigned)(tb - 3 * RQ_T + 1) * (unsigned)P + (unsigned)c0;
    }
    // ticks tb .. tb + 3 (tb a multiple of 4); true when the task is finished
    template <bool GUARD, bool FIX, int ORIENT, bool PUSH>
    __device__ __forceinline__ bool ticks4(int tb) {
        const bool st_lane = (lane >= RQ_HALO / 4 && lane < 32 - RQ_HALO / 4) && c0 < N;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int tau = tb + k;                                   // tau & 3 == k
            if (tau > t_last) return true;                            // every requested row (<= t_last + 1) has been taken
            // row tau + 1: out of the x ring
            await((k + 1) & 3, ph ^ (k == 3 ? 1u : 0u));
            take_x((k + 1) & 1, (k + 1) & 3);
#pragma unroll
            for (int i = 0; i < RQ_T; ++i) {
                const int r = tau - 3 * i;                            // R_i's row
                const int sr = (k + i) & 3;                           // r & 3
                const int par = (k + i) & 1;                          // R_i and B_i sweep the cells e = par, par + 2 of their rows
                const float4 zv = *reinterpret_cast<const float4 *>(zring + 512 * (r & 15));
                const float zn[4] = {zv.x, zv.y, zv.z, zv.w};
                if (i == 0) sweep_into<GUARD>(r, par, zn[par], zn[par + 2], XL[k & 1], S[0][(sr + 3) & 3], XL[(k + 1) & 1], S[0][sr]);
                else sweep_into<GUARD>(r, par, zn[par], zn[par + 2], S[i > 0 ? i - 1 : 0][sr], S[i][(sr + 3) & 3], S[i > 0 ? i - 1 : 0][(sr + 1) & 3], S[i][sr]);
                sweep<GUARD>(r - 1, par, Zc[i][0], Zc[i][1], S[i][(sr + 3) & 3], S[i][(sr + 2) & 3], S[i][sr]);              // B_i
                Zc[i][0] = zn[1 - par];                               // the next tick's B_i sweeps the other two cells of row r
                Zc[i][1] = zn[3 - par];
                if (FIX) fix_row<GUARD, ORIENT>(r - 2, S[i][(sr + 2) & 3], S[i][(sr + 1) & 3], S[i][(sr + 3) & 3]);          // F_i
            }
            const int yo = tau - 3 * RQ_T + 1;                        // this row is final now: F_3's
            if (yo >= ys && yo < ye) {
                const float(&o)[4] = S[RQ_T - 1][(k + 1) & 3];         // (k + 3 + 2) & 3
                float *outp = xout + ooff;
                if (!GUARD ? st_lane : (st_lane && c0 + 3 < N)) {
                    *reinterpret_cast<float4 *>(outp) = make_float4(o[0], o[1], o[2], o[3]);
                } else if (st_lane) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c0 + e < N) outp[e] = o[e];
                }
                if (PUSH) {
                    // the first / last RQ_VH rows of the slab are the neighbours' ghost rows of the next pass: written
                    // straight into their copy of the array over NVLink (whole quads: the pad columns are never read)
                    float *peer = (yo < push_lo) ? push_up : ((yo >= push_hi) ? push_dn : nullptr);
                    if (peer && st_lane) *reinterpret_cast<float4 *>(peer + ooff) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            // row tau + 5 replaces row tau - 11 (x0, codes) and row tau + 1 (x ring, taken in this tick; the shuffles
            // above have brought every lane past that read)
            if (FIX) __syncwarp();
            if (tau + 5 <= t_last + 1) {
                if (elect_one()) request<GUARD>(tau + 5, (k + 1) & 3);
            }
            next_row();
        }
        ph ^= 1u;
        return false;
    }
    template <bool GUARD, bool FIX, int ORIENT, bool PUSH>
    __device__ __forceinline__ bool block(int tb) {
#pragma unroll 1
        for (int q = 0; q < RQ_BLOCK / 4; ++q)
            if (ticks4<GUARD, FIX, ORIENT, PUSH>(tb + 4 * q)) return true;
        return false;
    }
};

// One task per CTA (= one warp).  The first n_edge * nsegs_e tasks are the strips that touch the left or right wall
// (strip 0 and the last n_edge - 1: mirror codes on every row and range tests, the slow ones) in short segments -- they
// go first --, the others the interior strips in segments of seg_rows.
// PUSH (row slabs): the tasks that produce the slab's first / last RQ_VH rows also write them into push_up / push_dn, the
// neighbours' copies of xout, so that only a barrier separates two passes.
template <bool PUSH>
__global__ void __launch_bounds__(RQ_THREADS, RQ_CTAS_PER_SM) k_rb_stream(const float *__restrict__ xin, float *__restrict__ xout,
                                                          const float *__restrict__ x0, const uint8_t *__restrict__ codes,
                                                          const uint8_t *__restrict__ chunk_flags, float a, float c_recip, int orient,
                                                          int row_lo, int row_hi, int nstrips, int n_edge, int nsegs, int seg_rows,
                                                          int nsegs_e, int seg_rows_e, float *push_up, float *push_dn,
                                                          const unsigned *__restrict__ run_if, EqLayout L) {
    EQ_DYN_SMEM(rq_smem);
    if (run_if && *run_if == 0u) return;                                   // the a == 0 shortcut was taken (k_a0_check)
    const int lane = (int)threadIdx.x;
    const int task = (int)blockIdx.x;
    const int ns = nstrips - n_edge;                                       // interior strips 1 .. ns
    if (task >= n_edge * nsegs_e + ns * nsegs) return;
    RqTask T;
    T.wsm = rq_smem;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b) mbar_init(smem_u32(T.wsm + RQ_B_OFF) + 16u * b, 1u);
    }
    __syncwarp();
    int sx, ys, ye;
    if (task < n_edge * nsegs_e) {
        const int e = task % n_edge;
        sx = e ? nstrips - e : 0;
        ys = row_lo + (task / n_edge) * seg_rows_e;
        ye = min(ys + seg_rows_e, row_hi);
    } else {
        const int t = task - n_edge * nsegs_e;
        sx = 1 + t % ns;
        ys = row_lo + (t / ns) * seg_rows;
        ye = min(ys + seg_rows, row_hi);
    }
    const int N = L.N;
    T.xin = xin; T.xout = xout; T.x0 = x0; T.codes = codes;
    T.push_up = push_up; T.push_dn = push_dn;
    T.push_lo = row_lo + RQ_VH; T.push_hi = row_hi - RQ_VH;
    T.a = a; T.c_recip = c_recip;
    T.X0c = sx * RQ_SW - RQ_HALO;
    T.ys = ys; T.ye = ye; T.lane = lane;
    T.init(L);
    const int X0c = T.X0c;
    const bool col_inside = X0c >= 1 && X0c + 127 <= N - 2;
    // does set_boundaries have anything to do in this task (away from rows 0 / N-1, which take the guarded ticks)?
    // AdjustRow / AdjustColumn: the (band, chunk) summaries of k_build_codes (band = (row-1)/32, chunk = column / EQ_LSX_CW);
    // Passive: only the frame carries codes
    const int NB = (N - 2 + 31) / 32, NC = (N + EQ_LSX_CW - 1) / EQ_LSX_CW;
    const int q_lo = max(X0c, 0) / EQ_LSX_CW, q_hi = min(X0c + 127, N - 1) / EQ_LSX_CW, nq = q_hi - q_lo + 1;   // nq <= 9
    const uint8_t *fl = chunk_flags + (orient == EQ_ADJUST_COLUMN ? (size_t)NB * NC : 0);
    // the flags of the rows that the F stages of block tb touch (tb - 11 .. tb + 13, one row of slack): <= 3 bands x 9 chunks
    auto block_flags = [&](int tb) -> unsigned {
        const int b_lo = max((max(tb - 12, 1) - 1) / 32, 0), b_hi = min((min(tb + 14, N - 2) - 1) / 32 + 1, NB - 1);
        const int total = max(b_hi - b_lo + 1, 0) * nq;
        return (lane < total) ? (unsigned)fl[(size_t)(b_lo + lane / nq) * NC + q_lo + lane % nq] : 0u;
    };
    bool task_codes = false;
    if (orient != EQ_PASSIVE && col_inside) {
        const int b_lo = max((max(ys - RQ_VH - 30, 1) - 1) / 32, 0), b_hi = min((min(ye + RQ_VH + 2, N - 2) - 1) / 32 + 1, NB - 1);
        const int total = max(b_hi - b_lo + 1, 0) * nq;
        int any = 0;
        for (int t = lane; t < total; t += 32) any |= fl[(size_t)(b_lo + t / nq) * NC + q_lo + t % nq];
        task_codes = __any_sync(0xffffffffu, any != 0) != 0;
    }
    const int t_first = ys - RQ_VH;
    int tb = (t_first >= 0) ? (t_first & ~3) : -((-t_first + 3) & ~3);
    // guarded ticks also run set_boundaries, so their rows need the codes
    T.with_codes = task_codes || !col_inside || tb - 12 < 1 || T.t_last + 21 > N - 2;
    const bool use_flags = task_codes;
    T.prologue(tb);
    unsigned nf = use_flags ? block_flags(tb) : 0u;
    for (;; tb += RQ_BLOCK) {
        const bool guard = !col_inside || tb - 12 < 1 || tb + 21 > N - 2;
        const bool fix = use_flags && __any_sync(0xffffffffu, nf != 0u);
        if (use_flags) nf = block_flags(tb + RQ_BLOCK);                    // needed RQ_BLOCK ticks from now
        bool fin;
        if (guard) {
            if (orient == EQ_ADJUST_ROW) fin = T.block<true, true, EQ_ADJUST_ROW, PUSH>(tb);
            else if (orient == EQ_ADJUST_COLUMN) fin = T.block<true, true, EQ_ADJUST_COLUMN, PUSH>(tb);
            else fin = T.block<true, true, EQ_PASSIVE, PUSH>(tb);
        } else if (fix) {
            if (orient == EQ_ADJUST_ROW) fin = T.block<false, true, EQ_ADJUST_ROW, PUSH>(tb);
            else fin = T.block<false, true, EQ_ADJUST_COLUMN, PUSH>(tb);   // (Passive: no codes away from the frame)
        } else {
            fin = T.block<false, false, EQ_PASSIVE, PUSH>(tb);
        }
        if (fin) break;
    }
}
