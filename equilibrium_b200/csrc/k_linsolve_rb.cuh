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
#include <type_traits>

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

// ---------------------------------------------------------------------------------------------
// k_rb_reg: the temporally blocked red-black solver with the tile held in REGISTERS.
//
// k_rb_tiled keeps its region in shared memory and pays 5 LDS + 1 STS (2-way bank conflicts: same-colour
// cells are 2 floats apart) per cell update: it is bound by the shared-memory issue rate, 5x above the HBM
// time of a pass.  Here a thread owns one column PAIR (an even and an odd column) over all RBR_H rows of the
// tile: 2 x 64 floats in registers.  For the cell of colour c in row y (element e = (c ^ y) & 1 of the pair,
// static once the row loop is unrolled: tile origins are even in both axes)
//     up / down            = the same element of rows y-1 / y+1      (own registers)
//     one horizontal side  = the other element of the pair           (own register)
//     the other side       = the neighbouring lane's other element   (one SHFL; across the 4 warps of a
//                            CTA through a 2 KB edge buffer in shared memory, lanes 0 / 31 only)
//     x0                   = a thread-private, conflict-free shared-memory slot (register spill space)
// so an update is 1 SHFL + 1 LDS + 6 FP.  set_boundaries runs on the same registers from a packed 4-bit
// direction code per cell (LEFT / RIGHT / UP / DOWN mirror; Passive's conditional frame copies are turned
// into the same codes with sign +), and only in tiles that hold a code at all.
//
// Tiles sit on a fixed lattice (output RBR_WO x RBR_HO cells, halo 3 cells per iteration on every side);
// passes ping-pong between two arrays exactly like k_rb_tiled.  Bit-identical to it and to the oracle's
// red-black restatement (ref_lin_solve_red_black): same expression tree, same sweep order.
// ---------------------------------------------------------------------------------------------
#define RBR_T RB_T
#define RBR_HALO RB_H                       // 12
#ifndef RBR_WARPS
#define RBR_WARPS 4                         // warps (64-column strips) per CTA
#endif
#define RBR_THREADS (32 * RBR_WARPS)
#define RBR_W (64 * RBR_WARPS)              // 256 columns per CTA
#define RBR_H 64                            // rows per CTA, all of them in registers
#define RBR_WO (RBR_W - 2 * RBR_HALO)       // 232
#define RBR_HO (RBR_H - 2 * RBR_HALO)       // 40
#define RBR_X0_BYTES (2u * RBR_H * RBR_THREADS * 4u)
#define RBR_EDGE_OFF RBR_X0_BYTES           // float edge[RBR_WARPS][2][RBR_H]
#define RBR_SMEM_BYTES (RBR_X0_BYTES + RBR_WARPS * 2u * RBR_H * 4u)

// RBR_INLINE_BARRIER = 1 (opt-in, not yet measured on GPUs): no kernel at all between two launches of a row-slab
// solve.  The tiles that read ghost rows wait, at their start, for the neighbour's "launch e-1 finished" flag; the last
// CTA of a launch to finish raises that flag in both neighbours.  Interior tiles never wait, so the neighbour barrier
// hides behind them.
#ifndef RBR_INLINE_BARRIER
#define RBR_INLINE_BARRIER 0
#endif
#if RBR_INLINE_BARRIER
#include "k_multigpu.cuh"
struct RbrSync {
    unsigned *mine;        // my sync words: [16] raised by the rank above, [17] by the rank below, [20] CTA counter
    unsigned *up, *down;   // the neighbours' sync words (nullptr at the ends / on one GPU)
    unsigned epoch;        // number of this launch (the same on every rank)
    unsigned wait_epoch;   // ghost rows are valid once the neighbours finished this launch; 0 = no wait (copied rows)
    int *error;
};
#define RBR_SYNC_PARAM , RbrSync sy
// ghost rows land in my memory while this kernel is already resident: read x through L2 (ld.global.cg), not through
// the non-coherent path the compiler picks for a const __restrict__ pointer
#define RBR_LD_X2(p) __ldcg(reinterpret_cast<const float2 *>(p))
#define RBR_LD_X1(p) __ldcg(p)
#else
#define RBR_SYNC_PARAM
#define RBR_LD_X2(p) (*reinterpret_cast<const float2 *>(p))
#define RBR_LD_X1(p) (*(p))
#endif

#define RBR_DIR_LEFT 1u
#define RBR_DIR_RIGHT 2u
#define RBR_DIR_UP 3u
#define RBR_DIR_DOWN 4u

__global__ void __launch_bounds__(RBR_THREADS, 8 / RBR_WARPS) k_rb_reg(const float *__restrict__ xin, float *__restrict__ xout,
                                                           const float *__restrict__ x0, const uint8_t *__restrict__ codes,
                                                           const uint8_t *__restrict__ chunk_flags,
                                                           const uint8_t *__restrict__ row_fluid,
                                                           const uint8_t *__restrict__ col_fluid, float a, float c_recip,
                                                           int orient, int iters, int row_lo, int row_hi, int tile_y0,
                                                           float *__restrict__ peer_up_out, float *__restrict__ peer_down_out,
                                                           const unsigned *__restrict__ run_if, EqLayout L RBR_SYNC_PARAM) {
    if (run_if && *run_if == 0u) return;                                  // the a == 0 shortcut was taken (k_a0_check)
    EQ_DYN_SMEM(rbr_smem);
    float *x0s = reinterpret_cast<float *>(rbr_smem);                    // [2][RBR_H][RBR_THREADS], thread-private slots
    float *edge = reinterpret_cast<float *>(rbr_smem + RBR_EDGE_OFF);     // [warp][side][row]
    const int N = L.N, P = L.P;
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gx0 = (int)blockIdx.x * RBR_WO - RBR_HALO;                 // even
    const int gy0 = (tile_y0 + (int)blockIdx.y) * RBR_HO - RBR_HALO;     // even
    const int lx = 64 * w + 2 * lane, gx = gx0 + lx;                     // my even column
    const bool in0 = (gx >= 0 && gx < N), in1 = (gx + 1 >= 0 && gx + 1 < N);
    const bool colok0 = (gx >= 1 && gx <= N - 2), colok1 = (gx + 1 >= 1 && gx + 1 <= N - 2);

#if RBR_INLINE_BARRIER
    if (sy.wait_epoch) {
        // tiles whose region reaches into the neighbours' rows read ghost rows the neighbours' previous launch pushed
        const bool reads_up = sy.up && gy0 < row_lo, reads_down = sy.down && gy0 + RBR_H > row_hi;
        if (reads_up || reads_down) {
            if (tid == 0) {
                if (reads_up) eq_xgpu_wait(sy.mine + 16, sy.wait_epoch, sy.error);
                if (reads_down) eq_xgpu_wait(sy.mine + 17, sy.wait_epoch, sy.error);
            }
            __syncthreads();
        }
    }
#endif
    float v[RBR_H][2];
    // ---- stage: x into registers, x0 into my shared-memory slots.  x0 goes global -> shared with cp.async (no
    // register in between: the 249 registers of this kernel leave none to batch loads in), so all of a thread's
    // 64 LDG.64 of x and 128 cp.async of x0 are in flight together.
    const uint32_t x0s_u32 = smem_u32(x0s);
    if (gx0 >= 0 && gx0 + RBR_W <= N && gy0 >= 0 && gy0 + RBR_H <= N) {   // the whole tile lies inside the grid
        const size_t o0 = (size_t)gy0 * P + gx;
#pragma unroll
        for (int y = 0; y < RBR_H; ++y) {
            const size_t o = o0 + (size_t)y * P;
            cp_async_4s(x0s_u32 + 4u * ((0 * RBR_H + y) * RBR_THREADS + tid), x0 + o);
            cp_async_4s(x0s_u32 + 4u * ((1 * RBR_H + y) * RBR_THREADS + tid), x0 + o + 1);
        }
#pragma unroll
        for (int y = 0; y < RBR_H; ++y) {
            const float2 xv = RBR_LD_X2(xin + o0 + (size_t)y * P);
            v[y][0] = xv.x;
            v[y][1] = xv.y;
        }
    } else {
#pragma unroll
        for (int y = 0; y < RBR_H; ++y) {
            const int gy = gy0 + y;
            const bool rowin = (gy >= 0 && gy < N);
            const size_t o = (size_t)gy * P + gx;                        // only dereferenced where rowin && in0 / in1
            v[y][0] = (rowin && in0) ? RBR_LD_X1(xin + o) : 0.f;
            v[y][1] = (rowin && in1) ? RBR_LD_X1(xin + o + 1) : 0.f;
            if (rowin && in0) cp_async_4s(x0s_u32 + 4u * ((0 * RBR_H + y) * RBR_THREADS + tid), x0 + o);
            else x0s[(0 * RBR_H + y) * RBR_THREADS + tid] = 0.f;
            if (rowin && in1) cp_async_4s(x0s_u32 + 4u * ((1 * RBR_H + y) * RBR_THREADS + tid), x0 + o + 1);
            else x0s[(1 * RBR_H + y) * RBR_THREADS + tid] = 0.f;
        }
    }
    cp_async_commit();
    // ---- does set_boundaries have anything to do in this tile? ----
    bool need_fix;
    if (orient == EQ_PASSIVE) {
        need_fix = (gx0 <= 0) || (gy0 <= 0) || (gx0 + RBR_W >= N) || (gy0 + RBR_H >= N);
    } else {
        // (band, chunk) summaries of k_build_codes: band = (row-1)/32, chunk = column / EQ_LSX_CW; a superset is fine
        const int NB = (N - 2 + 31) / 32, NC = (N + EQ_LSX_CW - 1) / EQ_LSX_CW;
        const int b_lo = max((max(gy0, 1) - 1) / 32 - 1, 0), b_hi = min((min(gy0 + RBR_H - 1, N - 2) - 1) / 32 + 1, NB - 1);
        const int q_lo = max(gx0, 0) / EQ_LSX_CW, q_hi = min(gx0 + RBR_W - 1, N - 1) / EQ_LSX_CW;
        const uint8_t *fl = chunk_flags + (orient == EQ_ADJUST_COLUMN ? (size_t)NB * NC : 0);
        const int nq = q_hi - q_lo + 1, total = max(b_hi - b_lo + 1, 0) * max(nq, 0);
        int any = 0;
        for (int t = tid; t < total; t += RBR_THREADS) any |= fl[(size_t)(b_lo + t / nq) * NC + q_lo + t % nq];
        need_fix = __syncthreads_or(any) != 0;
    }
    // packed direction codes: 4 bits per cell, cell (y, e) at bits 8*(y&3) + 4*e of cw[y>>2]
    uint32_t cw[RBR_H / 4];
#pragma unroll
    for (int i = 0; i < RBR_H / 4; ++i) cw[i] = 0u;
    if (need_fix) {
        // branch-free on purpose: every load below has a valid (clamped) address and is unconditional, so the 64
        // loads of a thread are in flight together; the predicates only mask the results
        const int gxc = in0 ? gx : 0;                                      // even, and column gxc + 1 <= P - 1 exists
        if (orient == EQ_PASSIVE) {
            // frame cells copy their interior neighbour (fluid.rs:182-186, conditional per quirk Q6)
            const bool cf0 = in0 && col_fluid[gxc] != 0, cf1 = in1 && col_fluid[min(gxc + 1, N - 1)] != 0;
            const bool fx0 = in0 && (gx == 0 || gx == N - 1), fx1 = in1 && (gx + 1 == 0 || gx + 1 == N - 1);
#pragma unroll
            for (int y = 0; y < RBR_H; ++y) {
                const int gy = gy0 + y, gyc = min(max(gy, 0), N - 1);
                const bool rowin = (gy == gyc), fy = (gy == 0 || gy == N - 1);
                const bool rf = row_fluid[gyc] != 0;
                unsigned d0 = 0u, d1 = 0u;
                if (fy) {
                    d0 = (in0 && !fx0 && cf0) ? (gy == 0 ? RBR_DIR_DOWN : RBR_DIR_UP) : 0u;
                    d1 = (in1 && !fx1 && cf1) ? (gy == 0 ? RBR_DIR_DOWN : RBR_DIR_UP) : 0u;
                } else {
                    d0 = (fx0 && rf) ? (gx == 0 ? RBR_DIR_RIGHT : RBR_DIR_LEFT) : 0u;
                    d1 = (fx1 && rf) ? (gx + 1 == 0 ? RBR_DIR_RIGHT : RBR_DIR_LEFT) : 0u;
                }
                if (!rowin) d0 = d1 = 0u;
                cw[y >> 2] |= (d0 | (d1 << 4)) << (8 * (y & 3));
            }
        } else {
            const unsigned shift = (orient == EQ_ADJUST_ROW) ? 0u : 2u;   // bits 0-1: 1 LEFT, 2 RIGHT; bits 2-3: 1 UP, 2 DOWN
            const unsigned bump = (orient == EQ_ADJUST_ROW) ? 0u : 2u;    // direction numbers: LEFT 1, RIGHT 2, UP 3, DOWN 4
#pragma unroll
            for (int y = 0; y < RBR_H; ++y) {
                const int gy = gy0 + y, gyc = min(max(gy, 0), N - 1);
                const bool rowin = (gy == gyc);
                const unsigned pair = *reinterpret_cast<const uint16_t *>(codes + (size_t)gyc * P + gxc);   // one 16-bit load
                unsigned d0 = ((pair & 255u) >> shift) & 3u, d1 = ((pair >> 8) >> shift) & 3u;
                d0 = (d0 && rowin && in0) ? d0 + bump : 0u;
                d1 = (d1 && rowin && in1) ? d1 + bump : 0u;
                cw[y >> 2] |= (d0 | (d1 << 4)) << (8 * (y & 3));
            }
        }
    }
    // interior rows of the grid inside this tile: y in [ylo, yhi] (uniform over the CTA; compared against the
    // unrolled row index, so the guarded sweep needs no precomputed per-row predicate)
    const int ylo = max(1 - gy0, 1), yhi = min(N - 2 - gy0, RBR_H - 2);
    const bool guard = !(gx0 >= 1 && gx0 + RBR_W - 1 <= N - 2 && gy0 >= 1 && gy0 + RBR_H - 1 <= N - 2);
    float *my_edge_l = edge + (w * 2 + 0) * RBR_H;                         // lane 0 publishes its even column here
    float *my_edge_r = edge + (w * 2 + 1) * RBR_H;                         // lane 31 its odd column
    const float *nb_edge_l = edge + ((max(w, 1) - 1) * 2 + 1) * RBR_H;     // right edge of the warp to my left
    const float *nb_edge_r = edge + (min(w + 1, RBR_WARPS - 1) * 2 + 0) * RBR_H;   // left edge of the warp to my right
    const bool has_l = (w > 0), has_r = (w < RBR_WARPS - 1);

    // lanes 0 and 31 publish their outer column: one predicated store per row (two separate `if`s compile to branches)
    float *const my_edge = (lane == 0) ? my_edge_l : my_edge_r;
    const bool is_edge_lane = (lane == 0 || lane == 31);
#pragma unroll
    for (int y = 0; y < RBR_H; ++y) {
        const float ev = (lane == 0) ? v[y][0] : v[y][1];
        if (is_edge_lane) my_edge[y] = ev;
    }
    cp_async_wait<0>();                                                    // my x0 slots have landed (thread-private)
    __syncthreads();

    // one half-sweep: the cells of colour c.  Everything a cell reads has the other colour, so the 62 updates of
    // a thread are independent; the loop only LOADS from shared memory (x0 slots, the neighbour warps' edge
    // columns) -- the edge columns of this colour are published after it, so that no store orders the rows.
    auto sweep = [&](auto colour_c, auto guard_c) {
        constexpr int C = decltype(colour_c)::value;
        constexpr bool GUARD = decltype(guard_c)::value;
#pragma unroll
        for (int y = 1; y <= RBR_H - 2; ++y) {
            const int e = (C ^ y) & 1;
            float left, right;
            if (e == 0) {
                float hn = __shfl_up_sync(0xffffffffu, v[y][1], 1);
                if (lane == 0 && has_l) hn = nb_edge_l[y];
                left = hn;
                right = v[y][1];
            } else {
                float hn = __shfl_down_sync(0xffffffffu, v[y][0], 1);
                if (lane == 31 && has_r) hn = nb_edge_r[y];
                left = v[y][0];
                right = hn;
            }
            const float nv = gs_update(x0s[(e * RBR_H + y) * RBR_THREADS + tid], right, left, v[y + 1][e], v[y - 1][e], a, c_recip);
            if (GUARD) {
                const bool ok = (e == 0 ? colok0 : colok1) && y >= ylo && y <= yhi;
                v[y][e] = ok ? nv : v[y][e];
            } else {
                v[y][e] = nv;
            }
        }
#pragma unroll
        for (int y = 1; y <= RBR_H - 2; ++y) {
            if (((C ^ y) & 1) == 0) { if (lane == 0) my_edge_l[y] = v[y][0]; }
            else { if (lane == 31) my_edge_r[y] = v[y][1]; }
        }
    };

    // set_boundaries (fluid.rs:252-272) on the registers.  Sources (wall cells; interior cells for Passive) are
    // never destinations, so the order inside the pass does not matter.  Rows in which no lane of the warp holds
    // a code are skipped with one vote.
    auto fixup = [&](auto orient_c) {
        constexpr int O = decltype(orient_c)::value;
        const bool edge_l = (lane == 0 && !has_l), edge_r = (lane == 31 && !has_r);
#pragma unroll
        for (int y = 0; y < RBR_H; ++y) {
            const unsigned byte = (cw[y >> 2] >> (8 * (y & 3))) & 0xffu;
            const float o0 = v[y][0], o1 = v[y][1];
            float ln = 0.f, rn = 0.f;
            if (O != EQ_ADJUST_COLUMN) {
                // the shuffles stay outside the vote branch: the compiler cannot know the branch is warp-uniform and
                // would route them through the WARPSYNC.COLLECTIVE slow path
                ln = __shfl_up_sync(0xffffffffu, o1, 1);                  // left neighbour of my even cell
                rn = __shfl_down_sync(0xffffffffu, o0, 1);                // right neighbour of my odd cell
            }
            if (!__any_sync(0xffffffffu, byte != 0u)) continue;
            const unsigned d0 = byte & 15u, d1 = byte >> 4;
            float n0 = o0, n1 = o1;
            if (O != EQ_ADJUST_COLUMN) {
                if (lane == 0 && has_l) ln = nb_edge_l[y];
                if (lane == 31 && has_r) rn = nb_edge_r[y];
                const float sl0 = (O == EQ_PASSIVE) ? ln : -ln, sr0 = (O == EQ_PASSIVE) ? o1 : -o1;
                const float sl1 = (O == EQ_PASSIVE) ? o0 : -o0, sr1 = (O == EQ_PASSIVE) ? rn : -rn;
                n0 = (d0 == RBR_DIR_LEFT && !edge_l) ? sl0 : ((d0 == RBR_DIR_RIGHT) ? sr0 : n0);
                n1 = (d1 == RBR_DIR_LEFT) ? sl1 : ((d1 == RBR_DIR_RIGHT && !edge_r) ? sr1 : n1);
            }
            if (O != EQ_ADJUST_ROW) {
                if (y > 0) {
                    const float u0 = v[y - 1][0], u1 = v[y - 1][1];
                    n0 = (d0 == RBR_DIR_UP) ? ((O == EQ_PASSIVE) ? u0 : -u0) : n0;
                    n1 = (d1 == RBR_DIR_UP) ? ((O == EQ_PASSIVE) ? u1 : -u1) : n1;
                }
                if (y < RBR_H - 1) {
                    const float b0 = v[y + 1][0], b1 = v[y + 1][1];
                    n0 = (d0 == RBR_DIR_DOWN) ? ((O == EQ_PASSIVE) ? b0 : -b0) : n0;
                    n1 = (d1 == RBR_DIR_DOWN) ? ((O == EQ_PASSIVE) ? b1 : -b1) : n1;
                }
            }
            v[y][0] = n0;
            v[y][1] = n1;
        }
    };

    for (int it = 0; it < iters; ++it) {
        if (guard) sweep(std::integral_constant<int, 0>{}, std::true_type{});
        else sweep(std::integral_constant<int, 0>{}, std::false_type{});
        __syncthreads();
        if (guard) sweep(std::integral_constant<int, 1>{}, std::true_type{});
        else sweep(std::integral_constant<int, 1>{}, std::false_type{});
        __syncthreads();
        if (need_fix) {
            if (orient == EQ_ADJUST_ROW) fixup(std::integral_constant<int, EQ_ADJUST_ROW>{});
            else if (orient == EQ_ADJUST_COLUMN) fixup(std::integral_constant<int, EQ_ADJUST_COLUMN>{});
            else fixup(std::integral_constant<int, EQ_PASSIVE>{});
            if (it + 1 < iters) {
                // the neighbouring warps may still be reading my edge columns inside their fixup (the values they mirror are
                // never changed by the pass, so the overlap was benign -- compute-sanitizer racecheck rightly flags it)
                __syncthreads();
#pragma unroll
                for (int y = 0; y < RBR_H; ++y) {
                    const float ev = (lane == 0) ? v[y][0] : v[y][1];
                    if (is_edge_lane) my_edge[y] = ev;
                }
            }
            __syncthreads();
        }
    }
    // ---- write the centre ----
    if (lx >= RBR_HALO && lx < RBR_W - RBR_HALO && in0) {
#pragma unroll
        for (int y = RBR_HALO; y < RBR_H - RBR_HALO; ++y) {
            const int gy = gy0 + y;
            if (gy >= row_lo && gy < row_hi && gy < N) {
                const size_t o = (size_t)gy * P + gx;
                if (in1) *reinterpret_cast<float2 *>(xout + o) = make_float2(v[y][0], v[y][1]);
                else xout[o] = v[y][0];
            }
        }
    }
#ifndef RBR_NO_PUSH
    // Row slabs: my first / last RBR_HALO rows are the neighbours' ghost rows for the next launch.  The tiles that
    // produce them write them straight into the neighbours' copy of xout over NVLink, so that the exchange between
    // two launches is only a barrier.  A separate, block-uniform branch: folded into the write-out loop above, the
    // extra selects cost the single-GPU kernel 15 % (11.1 -> 13.0 ms per 16384^2 solve).
    const int gy_lo = gy0 + RBR_HALO, gy_hi = gy0 + RBR_H - RBR_HALO;      // output rows of this tile
    const bool push_up = peer_up_out && gy_lo < row_lo + RBR_HALO && gy_hi > row_lo;
    const bool push_down = peer_down_out && gy_hi > row_hi - RBR_HALO && gy_lo < row_hi;
    if ((push_up || push_down) && lx >= RBR_HALO && lx < RBR_W - RBR_HALO && in0) {
#pragma unroll
        for (int y = RBR_HALO; y < RBR_H - RBR_HALO; ++y) {
            const int gy = gy0 + y;
            if (gy < row_lo || gy >= row_hi || gy >= N) continue;
            float *peer = (push_up && gy < row_lo + RBR_HALO) ? peer_up_out : ((push_down && gy >= row_hi - RBR_HALO) ? peer_down_out : nullptr);
            if (peer) {
                const size_t o = (size_t)gy * P + gx;
                if (in1) *reinterpret_cast<float2 *>(peer + o) = make_float2(v[y][0], v[y][1]);
                else peer[o] = v[y][0];
            }
        }
    }
#endif
#if RBR_INLINE_BARRIER
    if (sy.up || sy.down) {
        __threadfence_system();                    // my stores and pushes, before I count as finished
        __syncthreads();
        if (tid == 0) {
            const unsigned total = gridDim.x * gridDim.y;
            if (atomicAdd(sy.mine + 20, 1u) == total - 1u) {   // the last CTA of the launch
                __threadfence_system();
                sy.mine[20] = 0u;                  // the next launch starts after this kernel (same stream)
                if (sy.up) st_release_sys_u32(sy.up + 17, sy.epoch);       // I am the rank below my upper neighbour
                if (sy.down) st_release_sys_u32(sy.down + 16, sy.epoch);   // ... and the rank above my lower one
            }
        }
    }
#endif
}
