// k_stencils.cuh -- the non-iterative kernels of Fluid::step (fluid.rs:437-524):
// mask-derived tables, set_boundaries (sparse), divergence, gradient subtract,
// semi-Lagrangian advect, point sources, init.
//
// Every floating-point expression uses the explicit round-to-nearest intrinsics
// in the reference's evaluation order, so nvcc can neither contract to FMA nor
// reassociate (rustc does neither).
#pragma once
#include "eq_common.cuh"

// ---------------------------------------------------------------------------
// Mask-derived tables.  set_boundaries (fluid.rs:252-272) visits every cell but
// only changes NoWall cells that touch a wall (AdjustRow/AdjustColumn) or frame
// cells (Passive).  We precompute, per cell, which neighbour it mirrors.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned eq_cell_code(const uint8_t *cells, int i, int j, int N, int P) {
    if (cells[i + (size_t)j * P]) return EQ_CODE_WALL;                 // fluid.rs:141-143
    const int il = max(i - 1, 0), ir = min(i + 1, N - 1);              // :147-148 (clamped)
    const int ju = max(j - 1, 0), jd = min(j + 1, N - 1);              // :145-146
    unsigned code = 0;
    if (cells[ir + (size_t)j * P]) code |= EQ_CODE_ROW_RIGHT;          // right assigned last => wins (:158-163)
    else if (cells[il + (size_t)j * P]) code |= EQ_CODE_ROW_LEFT;      // :152-157
    if (cells[i + (size_t)ju * P]) code |= EQ_CODE_COL_UP;             // up assigned last => wins (:172-177)
    else if (cells[i + (size_t)jd * P]) code |= EQ_CODE_COL_DOWN;      // :166-171
    return code;
}

// pass 0: codes + row/column "has fluid" flags + counts; pass 1: fill the lists
__global__ void k_build_codes(const uint8_t *__restrict__ cells, uint8_t *__restrict__ codes,
                              uint8_t *row_fluid, uint8_t *col_fluid, uint8_t *chunk_flags, uint8_t *chunk_flags_tb,
                              unsigned *counts,
                              uint2 *row_list, uint2 *col_list, int pass, EqLayout L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= L.N) return;
    const unsigned code = eq_cell_code(cells, i, j, L.N, L.P);
    const unsigned o = (unsigned)i + (unsigned)j * (unsigned)L.P;
    if (pass == 0) {
        codes[o] = (uint8_t)code;
        if (!(code & EQ_CODE_WALL)) {
            row_fluid[j] = 1;
            col_fluid[i] = 1;
        }
        // (band, chunk) summaries for the wavefront solver: band = 32 rows from row 1, chunk = EQ_LSX_CW columns
        const int NB = (L.N - 2 + 31) / 32, NC = (L.N + EQ_LSX_CW - 1) / EQ_LSX_CW;
        // the temporally blocked solver uses bands that move up 2 rows per iteration: band b reads the
        // codes of rows 32b-SK .. 32b+32 (SK = 2(T-1)), so a code in row j concerns bands (j-1)/32 .. (j+SK)/32
        const int SK = 2 * (TBX_T - 1), NBP = (L.N - 2 + SK + 31) / 32;
        // AdjustColumn: the NoWall cells of rows 1 and N-2 always carry UP / DOWN (they touch the frame) and the
        // solver's fast loop handles exactly that; the flag marks everything else -- a code in another row, or
        // a cell of those two rows that is a wall or mirrors the other way.
        bool col_complex = (code & 12u) != 0;
        if (i >= 1 && i <= L.N - 2) {
            if (j == 1) col_complex = (code & (12u | EQ_CODE_WALL)) != EQ_CODE_COL_UP;
            else if (j == L.N - 2) col_complex = (code & (12u | EQ_CODE_WALL)) != EQ_CODE_COL_DOWN;
        }
        if ((code & 3u) || col_complex) {
            for (int bb = (j - 1) / 32; bb <= min((j + SK) / 32, NBP - 1); ++bb) {
                if (code & 3u) chunk_flags_tb[(size_t)bb * NC + i / EQ_LSX_CW] = 1;
                if (col_complex) chunk_flags_tb[(size_t)NBP * NC + (size_t)bb * NC + i / EQ_LSX_CW] = 1;
            }
        }
        const bool owned = (j >= L.row0 && j < L.row1);   // the sparse lists drive stand-alone boundary passes
        if (code & 3u) {
            if (owned) atomicAdd(&counts[0], 1u);
            chunk_flags[(size_t)min((j - 1) / 32, NB - 1) * NC + i / EQ_LSX_CW] = 1;
        }
        if (code & 12u) {
            if (owned) atomicAdd(&counts[1], 1u);
            chunk_flags[(size_t)NB * NC + (size_t)min((j - 1) / 32, NB - 1) * NC + i / EQ_LSX_CW] = 1;
            if (j / 32 < NB)   // row j is also row j0-1 of the band below (cross-band DOWN patch)
                chunk_flags[(size_t)NB * NC + (size_t)(j / 32) * NC + i / EQ_LSX_CW] = 1;
        }
    } else {
        // Passive frame copies as mirror codes of the frame cells (bits 4-6; needs the flags of pass 0)
        unsigned pc = WF_C_NONE;
        if (j >= 1 && j <= L.N - 2 && row_fluid[j]) pc = (i == 0) ? WF_C_R : ((i == L.N - 1) ? WF_C_L : pc);
        if (i >= 1 && i <= L.N - 2 && col_fluid[i]) pc = (j == 0) ? WF_C_D : ((j == L.N - 1) ? WF_C_U : pc);
        if (pc) codes[o] = (uint8_t)(code | (pc << EQ_CODE_PASSIVE_SHIFT));
    }
    if (pass == 1 && j >= L.row0 && j < L.row1) {
        if (code & 3u) {
            const unsigned slot = atomicAdd(&counts[2], 1u);
            row_list[slot] = make_uint2(o, (code & 3u) == EQ_CODE_ROW_RIGHT ? o + 1u : o - 1u);
        }
        if (code & 12u) {
            const unsigned slot = atomicAdd(&counts[3], 1u);
            col_list[slot] = make_uint2(o, (code & 12u) == EQ_CODE_COL_UP ? o - (unsigned)L.P : o + (unsigned)L.P);
        }
    }
}

// 4 corners (fluid.rs:265-271) from the *current* frame values.
// (the top corners belong to the rank that owns row 0, the bottom ones to the owner of row N-1)
__device__ __forceinline__ void eq_corners(float *x, const EqLayout &L) {
    const int N = L.N, P = L.P;
    const size_t last = (size_t)(N - 1) * P;
    const size_t prev = (size_t)(N - 2) * P;
    if (L.row0 == 0) {
        x[0] = __fmul_rn(0.5f, __fadd_rn(x[1], x[P]));
        x[N - 1] = __fmul_rn(0.5f, __fadd_rn(x[N - 2], x[(size_t)P + N - 1]));
    }
    if (L.row1 == N) {
        x[last] = __fmul_rn(0.5f, __fadd_rn(x[last + 1], x[prev]));
        x[last + N - 1] = __fmul_rn(0.5f, __fadd_rn(x[last + N - 2], x[prev + N - 1]));
    }
}

// ---------------------------------------------------------------------------
// lin_solve with a == 0, c == 1 (diffuse with diffusion or viscosity 0, the reference's default for the density,
// configs.rs:50-60; fluid.rs:286-297 -> :315-320 becomes x = (x0 + 0 * s) * 1).
// 0 * s is +-0 for every finite s and x0 + (+-0) == x0 bit for bit unless x0 is -0.0 (then the sign of s decides), so all
// K sweeps leave x == x0 in the interior and only the LAST set_boundaries shows -- provided no neighbour sum ever
// overflows or meets a non-finite value (0 * inf = NaN).  Every value a sum can meet is an interior x0, a value of the
// initial x (the not-yet-updated neighbours of the first sweep, and the frame), or the negation of one of those, so the
// guard is: all of them finite with |v| <= 1e37 (four of them cannot overflow) and no interior x0 equal to -0.0.
// k_a0_check raises *flag when the guard fails; k_a0_apply copies x0 into the interior when it did not; the solver
// kernels are launched regardless and return at once when *flag == 0 (`run_if`).  No host round trip.  With row slabs
// every rank checks the rows it owns and the flags are OR-ed over the ranks (k_flag_or_all) before anybody acts on them.
// ---------------------------------------------------------------------------
__global__ void k_a0_check(const float *__restrict__ x, const float *__restrict__ x0, unsigned *flag, EqLayout L) {
    const int N = L.N, P = L.P;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) * 4;     // 4 columns per thread (float4: P is a multiple of 32)
    if (g >= N) return;
    bool bad = false;
    for (int j = L.row0 + blockIdx.y; j < L.row1; j += gridDim.y) {   // the rows this rank owns (all of them on one GPU)
        const float4 a = *reinterpret_cast<const float4 *>(x + (size_t)j * P + g);
        const float4 b = *reinterpret_cast<const float4 *>(x0 + (size_t)j * P + g);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = g + k;
            if (i >= N) continue;
            bad = bad || !(fabsf(av[k]) <= 1e37f);                                  // also catches NaN
            const bool interior = (i >= 1 && i <= N - 2 && j >= 1 && j <= N - 2);
            if (interior) bad = bad || !(fabsf(bv[k]) <= 1e37f) || __float_as_uint(bv[k]) == 0x80000000u;
        }
    }
    if (bad) *flag = 1u;
}

__global__ void k_a0_apply(float *__restrict__ x, const float *__restrict__ x0, const unsigned *__restrict__ flag, EqLayout L) {
    if (*flag != 0u) return;                                        // guard failed: the wavefront solver does the work
    const int N = L.N, P = L.P;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (g >= N) return;
    for (int j = max(L.row0, 1) + blockIdx.y; j <= min(L.row1 - 1, N - 2); j += gridDim.y) {
        float4 v = *reinterpret_cast<const float4 *>(x0 + (size_t)j * P + g);
        if (g >= 4 && g + 3 <= N - 2) {
            *reinterpret_cast<float4 *>(x + (size_t)j * P + g) = v;
        } else {                                                    // the groups that hold a frame or a pad column
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (g + k >= 1 && g + k <= N - 2) x[(size_t)j * P + g + k] = vv[k];
        }
    }
}

// copy rows [r0, r1) of src into dst unless *flag == 0 (the red-black ping-pong result is only valid when the solver ran;
// flag == nullptr: always).  src may be a peer's array (the all-gather of the replicated exact solve)
__global__ void k_copy_rows_if(float *__restrict__ dst, const float *__restrict__ src, const unsigned *__restrict__ flag, int r0, int r1, EqLayout L) {
    if (flag && *flag == 0u) return;
    const int P = L.P;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (g >= P) return;
    for (int j = r0 + blockIdx.y; j < r1; j += gridDim.y)
        *reinterpret_cast<float4 *>(dst + (size_t)j * P + g) = *reinterpret_cast<const float4 *>(src + (size_t)j * P + g);
}

__global__ void k_corners(float *x, EqLayout L) {
    if (blockIdx.x == 0 && threadIdx.x == 0) eq_corners(x, L);
}

// set_boundaries(AdjustRow | AdjustColumn): x[dst] = -x[src] for the listed
// cells.  Sources are wall cells (never a dst), so the pass is order-free.
// Frame cells are walls, so the corners can be done by the same launch.
__global__ void k_bnd_list(float *x, const uint2 *__restrict__ list, unsigned n, EqLayout L) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const uint2 e = list[t];
        x[e.x] = -x[e.y];
    }
    if (t == 0) eq_corners(x, L);
}

// set_boundaries(Passive) (fluid.rs:179-187 + quirk Q6) and the corners.  The
// corner thread evaluates the post-copy frame values itself (it only reads
// interior cells, or frame cells nobody writes), so one launch suffices.
__global__ void k_bnd_passive(float *x, const uint8_t *__restrict__ row_fluid,
                              const uint8_t *__restrict__ col_fluid, EqLayout L) {
    const int N = L.N, P = L.P;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;   // 1 .. N-2
    if (t >= 1 && t <= N - 2) {
        if (col_fluid[t]) {
            if (L.row0 == 0) x[t] = x[t + (size_t)P];
            if (L.row1 == N) x[t + (size_t)(N - 1) * P] = x[t + (size_t)(N - 2) * P];
        }
        if (row_fluid[t] && t >= L.row0 && t < L.row1) {
            x[(size_t)t * P] = x[(size_t)t * P + 1];
            x[(size_t)t * P + N - 1] = x[(size_t)t * P + N - 2];
        }
    }
    if (t == 0) {
        const size_t r1 = P, rl = (size_t)(N - 1) * P, rp = (size_t)(N - 2) * P;
        const bool c1 = col_fluid[1], cl = col_fluid[N - 2], w1 = row_fluid[1], wl = row_fluid[N - 2];
        const float a00 = c1 ? x[r1 + 1] : x[1];                 // x[1,0] after the copy
        const float b00 = w1 ? x[r1 + 1] : x[r1];                // x[0,1]
        const float a0l = c1 ? x[rp + 1] : x[rl + 1];            // x[1,N-1]
        const float b0l = wl ? x[rp + 1] : x[rp];                // x[0,N-2]
        const float al0 = cl ? x[r1 + N - 2] : x[N - 2];         // x[N-2,0]
        const float bl0 = w1 ? x[r1 + N - 2] : x[r1 + N - 1];    // x[N-1,1]
        const float all_ = cl ? x[rp + N - 2] : x[rl + N - 2];   // x[N-2,N-1]
        const float bll = wl ? x[rp + N - 2] : x[rp + N - 1];    // x[N-1,N-2]
        if (L.row0 == 0) {
            x[0] = __fmul_rn(0.5f, __fadd_rn(a00, b00));
            x[N - 1] = __fmul_rn(0.5f, __fadd_rn(al0, bl0));
        }
        if (L.row1 == N) {
            x[rl] = __fmul_rn(0.5f, __fadd_rn(a0l, b0l));
            x[rl + N - 1] = __fmul_rn(0.5f, __fadd_rn(all_, bll));
        }
    }
}

// ---------------------------------------------------------------------------
// project, part 1 (fluid.rs:339-349): divergence and p = 0 on the interior.
//
// HBM-bound streaming stencils (16 and 20 algorithmic bytes per cell).  A thread owns one
// 16-byte column group (4 cells) and walks EQ_ST_ROWS consecutive rows: every global access is a
// coalesced 128-bit load or store, and the rows above / below (vy for the divergence, p for the
// gradient) roll through registers, so each element is requested once per block row instead of
// three times.  The horizontal neighbours of the group's end cells are two scalar loads that hit
// the L1 lines the neighbouring lanes fetch.  The frame columns 0 / N-1 and the pad columns are
// never written: the two groups that contain them fall back to predicated scalar stores.
// ---------------------------------------------------------------------------
#define EQ_ST_ROWS 8        // rows per thread
#define EQ_ST_THREADS 128   // one block covers 512 columns x EQ_ST_ROWS rows

__device__ __forceinline__ float4 ldg_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void stg_f4(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
// store the cells of a column group that lie on interior columns 1 .. N-2
__device__ __forceinline__ void eq_store_group(float *dst, int i0, int N, const float4 &v) {
    if (i0 >= 1 && i0 + 3 <= N - 2) {
        stg_f4(dst, v);
    } else {
        if (i0 >= 1 && i0 <= N - 2) dst[0] = v.x;
        if (i0 + 1 >= 1 && i0 + 1 <= N - 2) dst[1] = v.y;
        if (i0 + 2 <= N - 2) dst[2] = v.z;
        if (i0 + 3 <= N - 2) dst[3] = v.w;
    }
}

__device__ __forceinline__ float eq_div_cell(float vx_r, float vx_l, float vy_d, float vy_u, float nf) {
    float t = __fsub_rn(vx_r, vx_l);                                   // :341-345
    t = __fadd_rn(t, vy_d);
    t = __fsub_rn(t, vy_u);
    return __fdiv_rn(__fmul_rn(-0.5f, t), nf);
}

__global__ void __launch_bounds__(EQ_ST_THREADS) k_divergence(const float *__restrict__ vx, const float *__restrict__ vy,
                                                              float *__restrict__ div, float *__restrict__ p, EqLayout L) {
    const int N = L.N, P = L.P;
    const int i0 = 4 * (blockIdx.x * EQ_ST_THREADS + threadIdx.x);
    const int jb = max(L.row0, 1) + blockIdx.y * EQ_ST_ROWS;           // owned interior rows
    const int je = min(jb + EQ_ST_ROWS, min(L.row1, N - 1));
    if (i0 > N - 2 || jb >= je) return;
    const float nf = (float)N;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    size_t o = (size_t)i0 + (size_t)jb * P;
    float4 up = ldg_f4(vy + o - P), mid = ldg_f4(vy + o);
    for (int j = jb; j < je; ++j, o += P) {
        const float4 dn = ldg_f4(vy + o + P);
        const float4 c = ldg_f4(vx + o);
        const float l = (i0 > 0) ? vx[o - 1] : 0.f;
        const float r = vx[o + 4];                                     // column i0+4 <= P: inside the allocation
        float4 d;
        d.x = eq_div_cell(c.y, l, dn.x, up.x, nf);
        d.y = eq_div_cell(c.z, c.x, dn.y, up.y, nf);
        d.z = eq_div_cell(c.w, c.y, dn.z, up.z, nf);
        d.w = eq_div_cell(r, c.z, dn.w, up.w, nf);
        eq_store_group(div + o, i0, N, d);
        eq_store_group(p + o, i0, N, zero);                            // :346
        up = mid;
        mid = dn;
    }
}

// project, part 2 (fluid.rs:364-371): subtract the pressure gradient.
__device__ __forceinline__ float eq_grad_cell(float v, float p_hi, float p_lo, float nf) {
    return __fsub_rn(v, __fmul_rn(__fmul_rn(0.5f, __fsub_rn(p_hi, p_lo)), nf));
}

__global__ void __launch_bounds__(EQ_ST_THREADS) k_gradient(float *__restrict__ vx, float *__restrict__ vy,
                                                            const float *__restrict__ p, EqLayout L) {
    const int N = L.N, P = L.P;
    const int i0 = 4 * (blockIdx.x * EQ_ST_THREADS + threadIdx.x);
    const int jb = max(L.row0, 1) + blockIdx.y * EQ_ST_ROWS;
    const int je = min(jb + EQ_ST_ROWS, min(L.row1, N - 1));
    if (i0 > N - 2 || jb >= je) return;
    const float nf = (float)N;
    size_t o = (size_t)i0 + (size_t)jb * P;
    float4 up = ldg_f4(p + o - P), mid = ldg_f4(p + o);
    for (int j = jb; j < je; ++j, o += P) {
        const float4 dn = ldg_f4(p + o + P);
        const float l = (i0 > 0) ? p[o - 1] : 0.f;
        const float r = p[o + 4];
        const float4 a = ldg_f4(vx + o), b = ldg_f4(vy + o);
        float4 ax, by;
        ax.x = eq_grad_cell(a.x, mid.y, l, nf);
        ax.y = eq_grad_cell(a.y, mid.z, mid.x, nf);
        ax.z = eq_grad_cell(a.z, mid.w, mid.y, nf);
        ax.w = eq_grad_cell(a.w, r, mid.z, nf);
        by.x = eq_grad_cell(b.x, dn.x, up.x, nf);
        by.y = eq_grad_cell(b.y, dn.y, up.y, nf);
        by.z = eq_grad_cell(b.z, dn.z, up.z, nf);
        by.w = eq_grad_cell(b.w, dn.w, up.w, nf);
        eq_store_group(vx + o, i0, N, ax);
        eq_store_group(vy + o, i0, N, by);
        up = mid;
        mid = dn;
    }
}

// ---------------------------------------------------------------------------
// advect (fluid.rs:378-432).  One CTA per row because of the row-serial `break`
// (quirk Q4): the first cell f of row j whose back-traced sample leaves the grid
// copies its already-updated left neighbour and every cell after it keeps its
// stale destination value.  The flag depends on the velocity at the cell only.
//
// Single pass: the CTA walks the row in segments of EQ_ADV_SEG consecutive columns, EQ_ADV_U
// cells per thread (lane-interleaved, so velocity loads, the four bilinear gathers and the
// stores of a warp are unit-stride).  Per segment: back-trace, one block-wide vote on the
// flags, and only then the stores -- a segment without a flag (almost all of them) is written
// in full, the first segment with one is written up to f and ends the row.  The velocities of
// the next segment are requested before the vote, the EQ_ADV_U x 4 x NF gathers of a segment are
// independent loads in flight together.
// NF fields that share the velocity field (vx and vy self-advection,
// fluid.rs:469-489) are advected by one launch.
// ---------------------------------------------------------------------------
#define EQ_ADV_THREADS 256
#define EQ_ADV_U 4
#define EQ_ADV_SEG (EQ_ADV_THREADS * EQ_ADV_U)

struct AdvSample {
    float s0, s1, t0, t1;
    unsigned o00, o01, o10, o11;   // element offsets of (i0,j0) (i0,j1) (i1,j0) (i1,j1): 32 bits are enough up to 32768^2
    unsigned j0, j1;               // sample rows (the owner look-up of the row-slab path)
    bool flagged;
};

__device__ __forceinline__ float eq_clamp_rust(float v, float lo, float hi) {
    if (v < lo) v = lo;     // f32::clamp: NaN stays NaN
    if (v > hi) v = hi;
    return v;
}

// fi, fj = (float)i, (float)j (exact: the callers add small integers to one conversion per thread)
__device__ __forceinline__ AdvSample eq_backtrace(float fi, float fj, float u, float v, float dtx, float nf, int N, int P) {
    AdvSample r;
    float x = __fsub_rn(fi, __fmul_rn(dtx, u));                        // :400
    float y = __fsub_rn(fj, __fmul_rn(dtx, v));                        // :401 (delta_t_y == delta_t_x)
    x = eq_clamp_rust(x, 0.5f, __fsub_rn(nf, 1.0f));                   // :403
    y = eq_clamp_rust(y, 0.5f, __fsub_rn(nf, 1.0f));                   // :404
    const float i0 = floorf(x), i1 = __fadd_rn(i0, 1.0f);              // :406-407
    const float j0 = floorf(y), j1 = __fadd_rn(j0, 1.0f);              // :409-410
    r.s1 = __fsub_rn(x, i0);                                           // :412
    r.s0 = __fsub_rn(1.0f, r.s1);
    r.t1 = __fsub_rn(y, j0);
    r.t0 = __fsub_rn(1.0f, r.t1);
    // `as u32` saturates and maps NaN to 0 (:417-418); idx! then clamps to N-1
    const unsigned ui0 = min(__float2uint_rz(i0), (unsigned)(N - 1));
    const unsigned ui1 = min(__float2uint_rz(i1), (unsigned)(N - 1));
    r.j0 = min(__float2uint_rz(j0), (unsigned)(N - 1));
    r.j1 = min(__float2uint_rz(j1), (unsigned)(N - 1));
    const unsigned b0 = r.j0 * (unsigned)P, b1 = r.j1 * (unsigned)P;
    r.o00 = b0 + ui0;
    r.o01 = b1 + ui0;
    r.o10 = b0 + ui1;
    r.o11 = b1 + ui1;
    r.flagged = (i1 >= nf) || (j1 >= nf);                              // :420
    return r;
}

// The back-trace may leave the slab: each sample row is read from the rank that owns it (direct
// peer loads over NVLink; one rank => the local array).
template <bool PEERS>
__device__ __forceinline__ float eq_bilinear(const AdvSample &r, const EqPeerTable &t, int P, const float *mine, unsigned row0,
                                             unsigned row1) {
    // one rank: no owner look-up (it would index the parameter table dynamically and block the hoisting of the gathers).
    // Row slabs: almost every back-trace stays inside the rows this rank owns -- those lanes read the local array and
    // only the others search the table (a warp without such a lane skips the branch)
    const float *lo = mine, *hi = mine;
    if (PEERS && !(r.j0 >= row0 && r.j1 < row1)) {
        lo = eq_owner_base(t, r.j0);
        hi = eq_owner_base(t, r.j1);
    }
    const float a = lo[r.o00], b = hi[r.o01];
    const float c = lo[r.o10], d = hi[r.o11];
    const float l = __fadd_rn(__fmul_rn(r.t0, a), __fmul_rn(r.t1, b));
    const float h = __fadd_rn(__fmul_rn(r.t0, c), __fmul_rn(r.t1, d));
    return __fadd_rn(__fmul_rn(r.s0, l), __fmul_rn(r.s1, h));          // :424-428
}

template <int NF, bool PEERS>
__global__ void __launch_bounds__(EQ_ADV_THREADS) k_advect(float *__restrict__ dA, const EqPeerTable d0A,
                                                           float *__restrict__ dB, const EqPeerTable d0B,
                                                           const float *__restrict__ vx, const float *__restrict__ vy,
                                                           float dt, EqLayout L) {
    const int N = L.N, P = L.P;
    const int j = blockIdx.x + max(L.row0, 1);              // owned interior rows
    const float *mineA = PEERS ? d0A.base[L.rank] : d0A.base[0], *mineB = PEERS ? d0B.base[L.rank] : d0B.base[0];
    const float nf = (float)N;
    const float dtx = __fmul_rn(dt, (float)(N - 2));                   // :390
    const float fj = (float)j;
    const size_t row = (size_t)j * P;
    EQ_DYN_SMEM(adv_smem);
    int &s_first = *reinterpret_cast<int *>(adv_smem);
    if (threadIdx.x == 0) s_first = N;
    float u[EQ_ADV_U], v[EQ_ADV_U];
#pragma unroll
    for (int q = 0; q < EQ_ADV_U; ++q) {
        const int i = 1 + q * EQ_ADV_THREADS + (int)threadIdx.x;
        u[q] = (i <= N - 2) ? vx[row + i] : 0.f;
        v[q] = (i <= N - 2) ? vy[row + i] : 0.f;
    }
    for (int base = 1; base <= N - 2; base += EQ_ADV_SEG) {
        AdvSample r[EQ_ADV_U];
        int mine = N;
        const float fbase = (float)(base + (int)threadIdx.x);            // exact, as is fbase + q * 256 (< 2^24)
#pragma unroll
        for (int q = EQ_ADV_U - 1; q >= 0; --q) {
            const int i = base + q * EQ_ADV_THREADS + (int)threadIdx.x;
            r[q] = eq_backtrace(__fadd_rn(fbase, (float)(q * EQ_ADV_THREADS)), fj, u[q], v[q], dtx, nf, N, P);
            if (i <= N - 2 && r[q].flagged) mine = i;                  // descending q: the smallest flagged column
        }
        // velocities of the next segment (in flight across the vote and the gathers)
#pragma unroll
        for (int q = 0; q < EQ_ADV_U; ++q) {
            const int i = base + EQ_ADV_SEG + q * EQ_ADV_THREADS + (int)threadIdx.x;
            u[q] = (i <= N - 2) ? vx[row + i] : 0.f;
            v[q] = (i <= N - 2) ? vy[row + i] : 0.f;
        }
        const int any = __syncthreads_or(mine < N);                    // also orders the s_first initialisation
        int f = N;
        if (any) {
            if (mine < N) atomicMin(&s_first, mine);
            __syncthreads();
            f = s_first;                                               // first flagged column of the row
        }
        if (!any) {
            // the common case, branch-free: all EQ_ADV_U x 4 x NF gathers are issued before the first use (cells past
            // column N-2 back-trace to valid clamped addresses; only their stores are predicated)
            float a[EQ_ADV_U], b[EQ_ADV_U];
#pragma unroll
            for (int q = 0; q < EQ_ADV_U; ++q) {
                a[q] = eq_bilinear<PEERS>(r[q], d0A, P, mineA, (unsigned)L.row0, (unsigned)L.row1);
                b[q] = (NF == 2) ? eq_bilinear<PEERS>(r[q], d0B, P, mineB, (unsigned)L.row0, (unsigned)L.row1) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < EQ_ADV_U; ++q) {
                const int i = base + q * EQ_ADV_THREADS + (int)threadIdx.x;
                if (i <= N - 2) {
                    dA[row + i] = a[q];
                    if (NF == 2) dB[row + i] = b[q];
                }
            }
            continue;
        }
        // the segment that holds the first flagged column f of the row: cells before f as usual, cell f copies
        // its already-updated left neighbour, the rest of the row keeps its stale values
        const int last = min(f, N - 2);
#pragma unroll
        for (int q = 0; q < EQ_ADV_U; ++q) {
            const int i = base + q * EQ_ADV_THREADS + (int)threadIdx.x;
            if (i > last) continue;
            float a, b = 0.f;
            if (i < f) {
                a = eq_bilinear<PEERS>(r[q], d0A, P, mineA, (unsigned)L.row0, (unsigned)L.row1);
                if (NF == 2) b = eq_bilinear<PEERS>(r[q], d0B, P, mineB, (unsigned)L.row0, (unsigned)L.row1);
            } else if (i == 1) {                                       // :421, f == 1: the frame cell, untouched
                a = dA[row];
                if (NF == 2) b = dB[row];
            } else {                                                   // :421 copy of the updated left cell
                const AdvSample rl = eq_backtrace((float)(i - 1), fj, vx[row + i - 1], vy[row + i - 1], dtx, nf, N, P);
                a = eq_bilinear<PEERS>(rl, d0A, P, mineA, (unsigned)L.row0, (unsigned)L.row1);
                if (NF == 2) b = eq_bilinear<PEERS>(rl, d0B, P, mineB, (unsigned)L.row0, (unsigned)L.row1);
            }
            dA[row + i] = a;
            if (NF == 2) dB[row + i] = b;
        }
        break;                                                         // :422 -- the rest of the row keeps its stale values
    }
}

// ---------------------------------------------------------------------------
// sources and initial condition
// ---------------------------------------------------------------------------
// add_density (fluid.rs:120-124) / add_velocity (fluid.rs:127-131) as one stream-ordered launch
__global__ void k_add_source(float *density, float *scratch, float *vx, float *vy, unsigned o,
                             float dd, float dvx, float dvy, int what) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (what & 1) {
            density[o] = __fadd_rn(density[o], dd);
            scratch[o] = __fadd_rn(scratch[o], dd);
        }
        if (what & 2) {
            vx[o] = __fadd_rn(vx[o], dvx);
            vy[o] = __fadd_rn(vy[o], dvy);
        }
    }
}

// ---- device-side add_noise (fluid.rs:575-599) ----------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; the Random123 known-answer vectors are in tests/test_sources.py): counter =
// (frame, 0), key = seed.  Integer work only, so host restatements agree bit for bit.
__host__ __device__ __forceinline__ void eq_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = (uint32_t)p1;
        c[2] = n2;
        c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// One frame's impulse: a uniformly random grid point (:584-585) rotated about the centre (:587-593, the rotation of
// geo's rotate_around_point: x' = cos*(x-cx) - sin*(y-cy) + cx) and added, times `gain` (2.0, :595-596), to the velocity
// of the centre cell.  cos/sin come from the host: the angle depends on delta_t only (:578-583).
__global__ void k_add_noise(float *vx, float *vy, uint32_t seed_lo, uint32_t seed_hi, uint32_t frame_lo, uint32_t frame_hi,
                            float cos_t, float sin_t, float gain, EqLayout L) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t c[4] = {frame_lo, frame_hi, 0u, 0u};
    eq_philox4x32_10(c, seed_lo, seed_hi);
    const uint32_t N = (uint32_t)L.N;
    const uint32_t rx = (uint32_t)(((uint64_t)c[0] * N) >> 32), ry = (uint32_t)(((uint64_t)c[1] * N) >> 32);   // [0, N)
    const float ctr = (float)(N / 2u);
    const float dx = __fsub_rn((float)rx, ctr), dy = __fsub_rn((float)ry, ctr);
    const float px = __fadd_rn(__fsub_rn(__fmul_rn(cos_t, dx), __fmul_rn(sin_t, dy)), ctr);
    const float py = __fadd_rn(__fadd_rn(__fmul_rn(sin_t, dx), __fmul_rn(cos_t, dy)), ctr);
    const size_t o = (size_t)(N / 2u) + (size_t)(N / 2u) * (size_t)L.P;
    vx[o] = __fadd_rn(vx[o], __fmul_rn(px, gain));
    vy[o] = __fadd_rn(vy[o], __fmul_rn(py, gain));
}

// Dense source field a la Stam's add_source: x[i,j] += scale * s[i,j] on every cell of the grid (one read of each, one
// write: 12 B per cell).  One thread per float4 of a row; the pad columns beyond N are left alone.
__global__ void k_add_field(float *__restrict__ x, const float *__restrict__ s, float scale, EqLayout L) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;               // float4 index within the row
    const int N = L.N, i = q * 4;
    if (i >= N) return;
    for (int j = blockIdx.y; j < N; j += gridDim.y) {
        const size_t o = (size_t)j * L.P + i;
        if (i + 3 < N) {
            float4 a = *reinterpret_cast<const float4 *>(x + o);
            const float4 b = *reinterpret_cast<const float4 *>(s + o);
            a.x = __fadd_rn(a.x, __fmul_rn(scale, b.x));
            a.y = __fadd_rn(a.y, __fmul_rn(scale, b.y));
            a.z = __fadd_rn(a.z, __fmul_rn(scale, b.z));
            a.w = __fadd_rn(a.w, __fmul_rn(scale, b.w));
            *reinterpret_cast<float4 *>(x + o) = a;
        } else {
            for (int e = 0; i + e < N; ++e) x[o + e] = __fadd_rn(x[o + e], __fmul_rn(scale, s[o + e]));
        }
    }
}

// field[x,y] += amount on [x0,x1) x [y0,y1)   (init_velocities :542-548, init_density :534-538)
__global__ void k_add_rect(float *f, int x0, int y0, int x1, int y1, float amount, EqLayout L) {
    const int i = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = y0 + blockIdx.y;
    if (i < x1 && j < y1) {
        const size_t o = (size_t)i + (size_t)j * L.P;
        f[o] = __fadd_rn(f[o], amount);
    }
}

// cells[x,y] = value on [x0,x1) x [y0,y1) (init_walls :552-570, fill_obstacle :610-619; caller clamps)
__global__ void k_set_cells_rect(uint8_t *cells, int x0, int y0, int x1, int y1, uint8_t value, EqLayout L) {
    const int i = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = y0 + blockIdx.y;
    if (i < x1 && j < y1) cells[(size_t)i + (size_t)j * L.P] = value;
}

// sum over the interior of div^2 with the stencil of fluid.rs:341-345 (diagnostic only)
__global__ void k_divergence_sq(const float *__restrict__ vx, const float *__restrict__ vy, double *out,
                                EqLayout L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y + max(L.row0, 1);
    double v = 0.0;
    if (i >= 1 && i <= L.N - 2) {
        const size_t o = (size_t)i + (size_t)j * L.P;
        float t = __fsub_rn(vx[o + 1], vx[o - 1]);
        t = __fadd_rn(t, vy[o + L.P]);
        t = __fsub_rn(t, vy[o - L.P]);
        const float d = __fdiv_rn(__fmul_rn(-0.5f, t), (float)L.N);
        v = (double)d * (double)d;
    }
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out, v);
}

// ---------------------------------------------------------------------------
// density + cells_type -> RGBA, the pixel loop of RenderingListener::render_image
// (renderer_helpers.rs:145-167): wall => obstacle colour; else density != 0 =>
// [(density * fluid.r as f32) as u8, fluid.g, density as u8, 1]; else the world colour.
// `as u8` saturates and maps NaN to 0.  Output is compact (pitch N), rows row0.. of the slab.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned eq_f32_as_u8(float v) { return min(__float2uint_rz(v), 255u); }

__global__ void k_render_rgba(const float *__restrict__ density, const uint8_t *__restrict__ cells,
                              uint32_t *__restrict__ out, uint32_t world, uint32_t fluid, uint32_t obstacle, EqLayout L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y + L.row0;
    if (i >= L.N) return;
    const size_t o = (size_t)i + (size_t)j * L.P;
    const float d = density[o];
    uint32_t px = world;
    if (cells[o]) {
        px = obstacle;
    } else if (d != 0.0f) {
        const unsigned r = eq_f32_as_u8(__fmul_rn(d, (float)(fluid & 255u)));
        px = r | (fluid & 0x0000ff00u) | (eq_f32_as_u8(d) << 16) | (1u << 24);
    }
    out[(size_t)i + (size_t)(j - L.row0) * L.N] = px;
}
