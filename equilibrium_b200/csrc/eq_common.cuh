// eq_common.cuh -- shared device helpers and the HBM data layout.
//
// Layout of one field in HBM (DESIGN.md "Data layout"):
//   float buf[(N + EQ_ROW_PAD) * P],  P = round_up(N, 32)
//   cell (x, y) lives at buf[x + y * P]  (row-major like idx!, fluid.rs:31-35)
// Rows N..N+EQ_ROW_PAD-1 and columns N..P-1 are zero padding that no kernel
// ever writes a non-padding value to; they let the wavefront kernel stage
// 34-row x 32-column tiles with 16-byte cp.async and no bounds checks.
#pragma once
#include <stdint.h>

#include "../../include/equilibrium_cuda.h"

#ifdef EQ_HOST_EMU
// tests/emu: same sources compiled with g++ against a host SIMT emulator (tests only)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define EQ_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define EQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define EQ_ROW_PAD 48
#ifndef TBX_T
#define TBX_T 2      // iterations per job of the temporally blocked wavefront solver (k_linsolve_tb.cuh)
#endif
#ifndef EQ_LSX_CW
#define EQ_LSX_CW 16   // chunk width (columns) of the wavefront solver: staging / write-back / flag granularity
#endif

// cells_type / derived per-cell codes (one byte per cell, same pitch P):
//   bits 0-1  AdjustRow fix-up:    0 none, 1 take -x[i-1,j], 2 take -x[i+1,j]   (fluid.rs:151-164)
//   bits 2-3  AdjustColumn fix-up: 0 none, 1 take -x[i,j-1], 2 take -x[i,j+1]   (fluid.rs:165-178)
//   bit 7     the cell itself is DefaultWall
#define EQ_CODE_ROW_LEFT 1u
#define EQ_CODE_ROW_RIGHT 2u
#define EQ_CODE_COL_UP (1u << 2)
#define EQ_CODE_COL_DOWN (2u << 2)
#define EQ_CODE_WALL 0x80u
// bits 4-6, frame cells only: the Passive frame copies of fluid.rs:182-186 (quirk Q6) written as a mirror code of the
// FRAME cell, in the unified numbering the register wavefront solver uses for every orientation:
//   (0, j) takes x[1, j] = R, (N-1, j) takes x[N-2, j] = L   -- if row j holds a NoWall cell
//   (i, 0) takes x[i, 1] = D, (i, N-1) takes x[i, N-2] = U   -- if column i holds a NoWall cell
#define EQ_CODE_PASSIVE_SHIFT 4
#define WF_C_NONE 0u
#define WF_C_L 1u   // takes x[i-1, j]
#define WF_C_R 2u   // takes x[i+1, j]
#define WF_C_U 3u   // takes x[i, j-1]
#define WF_C_D 4u   // takes x[i, j+1]

struct EqLayout {
    int N;      // grid is N x N
    int P;      // row pitch in elements
    int rows;   // allocated rows = N + EQ_ROW_PAD
    // Row slab owned by this rank (multi-GPU).  Every rank allocates the full grid and
    // uses global coordinates; only rows [row0, row1) are authoritative here, rows row0-1 and row1
    // are ghost copies refreshed by k_halo_exchange.  Single GPU: row0 = 0, row1 = N.
    int row0, row1;
    int rank, world;
};

#define EQ_MAX_RANKS 8
// where each rank keeps a field, and which rows it owns (for gathers that leave the slab)
struct EqPeerTable {
    const float *base[EQ_MAX_RANKS];
    int row_begin[EQ_MAX_RANKS + 1];
    int world;
};
__device__ __forceinline__ const float *eq_owner_base(const EqPeerTable &t, unsigned j) {
    if (t.world <= 1) return t.base[0];
    int r = 0;
#pragma unroll
    for (int i = 1; i < EQ_MAX_RANKS; ++i)
        if (i < t.world && (int)j >= t.row_begin[i]) r = i;
    return t.base[r];
}

#ifndef EQ_HOST_EMU
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_s32(const int *p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte async copy global -> shared, L2 only (.cg): the data may have been
// written by another SM moments ago, so L1 must not serve it.
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// explicit 32-bit shared-memory addressing (keeps generic->shared conversions out of inner loops)
// The result is laundered through an asm so the compiler keeps it in a register instead of
// re-deriving the shared window base (S2R SR_CgaCtaId + LEA) at every predicated use.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u8(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_16s(uint32_t saddr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}
// 4-byte async copy global -> shared (through L1: read-only input of the kernel that issues it)
__device__ __forceinline__ void cp_async_4s(uint32_t saddr, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gmem) : "memory");
}
// named barrier over a subset of the CTA's warps (all `nthreads` threads must call it)
__device__ __forceinline__ void eq_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// pull a line into L2 ahead of time (no register, no shared memory)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// mbarrier (shared::cta) -- each barrier gets a 16-byte slot (8 used on the device)
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
// re-initialising a live mbarrier is undefined (PTX ISA, mbarrier.init): invalidate it first.  On B200 an init
// without the inval left the OLD phase in place -- a kernel whose barriers complete an odd number of phases per job
// then waits on the wrong parity in the next job of the same CTA.
__device__ __forceinline__ void mbar_inval(uint32_t a) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while)
__device__ __forceinline__ bool mbar_test_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
// arrive on the mbarrier once all cp.async of the executing thread issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t a) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
// bulk (TMA, 1-D) copy global -> shared: `bytes`, both addresses multiples of 16; completes `bytes` of the mbarrier's
// transaction count when the data has landed (SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t saddr, const void *gmem, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr),
                 "l"(gmem), "r"(bytes), "r"(mbar)
                 : "memory");
}
// true in exactly one lane of the (converged) warp: the compiler then issues what follows once, without a loop over lanes
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0u;
}
// one arrival that also announces `bytes` of pending bulk-copy traffic (issued before or after: the count is signed)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_release_cta_u32(uint32_t a, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire_cta_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
#endif  // !EQ_HOST_EMU

// Predicated stores as ONE instruction (an `if` around a store in the middle of an unrolled, shuffle-heavy loop makes the
// compiler treat the warp as possibly divergent from there on: every later shuffle becomes WARPSYNC.COLLECTIVE + SHFL).
#ifdef EQ_HOST_EMU
static inline void st_shared_f32_if(bool p, uint32_t a, float v) { if (p) sts_f32(a, v); }
static inline void st_global_f32_if(bool p, float *ptr, float v) { if (p) *ptr = v; }
#else
__device__ __forceinline__ void st_shared_f32_if(bool p, uint32_t a, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.f32 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void st_global_f32_if(bool p, float *ptr, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.f32 [%1], %2;\n\t}" ::"r"((unsigned)p), "l"(ptr), "f"(v) : "memory");
}
#endif

// The Gauss-Seidel update of fluid.rs:315-320 with the reference's expression
// tree: (x0 + a * (((right + left) + down) + up)) * c_recip, every operation
// individually rounded (no FMA: rustc never contracts).
__device__ __forceinline__ float gs_update(float x0, float right, float left, float down, float up,
                                           float a, float c_recip) {
    float s = __fadd_rn(right, left);
    s = __fadd_rn(s, down);
    s = __fadd_rn(s, up);
    return __fmul_rn(__fadd_rn(x0, __fmul_rn(a, s)), c_recip);
}

// Warp scheduler balance for the warp-specialised kernels.  A warp runs on sub-partition
// (hardware warp slot % 4); a 4-warp CTA occupies slots 4s..4s+3, so "warp 0 computes" would put
// the compute warps of ALL co-resident CTAs on sub-partition 0 while the other three idle along
// with the loaders and storers.  Rotating the roles by the CTA's slot number spreads them.
// Roles are still one warp each whatever the slots turn out to be: only balance depends on it.
__device__ __forceinline__ unsigned eq_cta_slot_rotation() {
#ifdef EQ_HOST_EMU
    return blockIdx.x & 3u;
#else
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    return (wid >> 2) & 3u;
#endif
}
