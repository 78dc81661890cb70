// k_multigpu.cuh -- row-slab decomposition across GPUs: halo rows and barriers.
//
// One process per GPU; every rank maps its neighbours' arrays (CUDA IPC, or plain peer access
// inside one process) and kernels read / write peer memory directly over NVLink: there is no
// host round trip and no library collective on the data path.  Cross-GPU ordering uses
// system-scope release/acquire on small flag arrays that live in the *consumer's* memory (the
// producer stores remotely, the consumer polls locally).
#pragma once
#include "eq_common.cuh"

#define EQ_SYNC_WORDS 32          // [0..3] halo slots (up A/B, down A/B), [8..15] all-rank barrier, [16..31] flag OR (2 x 8)
#ifdef EQ_HOST_EMU
#define EQ_XGPU_SPIN_LIMIT (1u << 30)   // emulated ranks are OS threads on a few cores: a spin is a yield, not 100 ns
#else
#define EQ_XGPU_SPIN_LIMIT (1u << 24)
#endif

struct EqHaloArgs {
    float *field;                 // my copy
    float *peer_up, *peer_down;   // the same field on rank-1 / rank+1 (nullptr at the ends)
    unsigned *sync;               // my slots, written by the neighbours
    unsigned *sync_up, *sync_down;// the neighbours' slots
    unsigned epoch;
    int nrows;                    // boundary rows pushed each way (1 for the stencils, 3T for tiled red-black)
    int *error;
    const unsigned *run_if;       // not null: skip when *run_if == 0 (exchanges that belong to a solve the a == 0 shortcut
                                  // replaced; the flag is the same on every rank, k_flag_or_all)
};

__device__ __forceinline__ bool eq_xgpu_wait(const unsigned *slot, unsigned epoch, int *error) {
    unsigned spins = 0;
    while (ld_relaxed_sys_u32(slot) < epoch) {
        __nanosleep(100);
        if ((++spins & 4095u) == 0) {
            if (spins >= EQ_XGPU_SPIN_LIMIT) {
                *error = 4;
                return false;
            }
            if (ld_volatile_s32(error) != 0) return false;
        }
    }
    (void)ld_acquire_sys_u32(slot);
    return true;
}

// Block 0 talks to rank-1, block 1 to rank+1.  Phase A: "everything I launched before this kernel
// (including my writes into your memory) is done" -- exchanged before any row is copied, because
// the neighbour's solver may still be patching my boundary row.  Phase B: push my boundary row
// into the neighbour's ghost row, then "pushed".
__global__ void __launch_bounds__(1024) k_halo_exchange(EqHaloArgs a, EqLayout L) {
    if (a.run_if && *a.run_if == 0u) return;
    const int dir = blockIdx.x;
    float *peer = dir == 0 ? a.peer_up : a.peer_down;
    if (!peer) return;
    unsigned *peer_sync = dir == 0 ? a.sync_up : a.sync_down;
    const unsigned *mine = a.sync + dir * 2;
    unsigned *theirs = peer_sync + (1 - dir) * 2;          // I am the neighbour's (1-dir) side
    EQ_DYN_SMEM(halo_smem);
    int &ok = *reinterpret_cast<int *>(halo_smem);
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys_u32(theirs + 0, a.epoch);
        ok = eq_xgpu_wait(mine + 0, a.epoch, a.error) ? 1 : 0;
    }
    __syncthreads();
    if (!ok || a.nrows <= 0) return;                       // nrows == 0: neighbour barrier only (the producer kernel
                                                           // pushed its boundary rows itself, see k_rb_reg)
    const int row = dir == 0 ? L.row0 : L.row1 - a.nrows;  // first / last owned rows
    const float4 *src = reinterpret_cast<const float4 *>(a.field + (size_t)row * L.P);
    float4 *dst = reinterpret_cast<float4 *>(peer + (size_t)row * L.P);
    for (int i = threadIdx.x; i < a.nrows * (L.P / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys_u32(theirs + 1, a.epoch);
        eq_xgpu_wait(mine + 1, a.epoch, a.error);
    }
}

struct EqBarrierArgs {
    unsigned *sync;                       // my slots
    unsigned *peer_sync[EQ_MAX_RANKS];    // everybody's slots
    int rank, world;
    unsigned epoch;
    int *error;
};

// OR of one flag word over all ranks (the guard of the a == 0 lin_solve shortcut must come out the same everywhere).
// Rank r stores (epoch << 1 | flag) into slot r of every peer, then waits for every peer's word of this epoch.  Two
// sets of slots alternate with the epoch: a peer can be at most one reduction ahead of me.
__global__ void k_flag_or_all(EqBarrierArgs a, unsigned *flag) {
    const int i = threadIdx.x;
    const unsigned base = 16u + 8u * (a.epoch & 1u);
    unsigned mine = *flag != 0u ? 1u : 0u;
    if (i < a.world && i != a.rank) {
        __threadfence_system();
        st_release_sys_u32(a.peer_sync[i] + base + a.rank, (a.epoch << 1) | mine);
        const unsigned *slot = a.sync + base + i;
        unsigned spins = 0, v;
        while (((v = ld_relaxed_sys_u32(slot)) >> 1) < a.epoch) {
            __nanosleep(100);
            if ((++spins & 4095u) == 0 && (spins >= EQ_XGPU_SPIN_LIMIT || ld_volatile_s32(a.error) != 0)) {
                if (spins >= EQ_XGPU_SPIN_LIMIT) *a.error = 4;
                break;
            }
        }
        v = ld_acquire_sys_u32(slot);
        if (v & 1u) *flag = 1u;                               // (several threads may store the same 1)
    }
}

// All ranks: "everything I launched before this kernel is done" (needed around advect, whose
// back-trace may read any rank's rows).
__global__ void k_barrier_all(EqBarrierArgs a) {
    const int i = threadIdx.x;
    if (i >= a.world || i == a.rank) return;
    __threadfence_system();
    st_release_sys_u32(a.peer_sync[i] + 8 + a.rank, a.epoch);
    eq_xgpu_wait(a.sync + 8 + i, a.epoch, a.error);
}
