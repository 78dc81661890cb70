// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE (see cuda_emu.h).
#include "cuda_emu.h"

#include <chrono>

namespace eq_emu {
thread_local Ctx ctx;

Mode mode_for(const char *name) {
    if (strstr(name, "k_linsolve_exact") || strstr(name, "k_halo_exchange")) return CONCURRENT_GRID;
    if (strstr(name, "k_advect") || strstr(name, "k_divergence_sq")) return BLOCK_THREADS;
    return SEQUENTIAL;
}

static void run_block_threads(BlockState &bs, dim3 grid, dim3 block, uint3_emu bid,
                              const std::function<void()> &body, std::vector<std::thread> &pool) {
    const int nthreads = (int)(block.x * block.y * block.z);
    for (int t = 0; t < nthreads; ++t) {
        BlockState *bsp = &bs;
        const std::function<void()> *bodyp = &body;
        pool.emplace_back([bsp, bodyp, grid, block, t, bid]() {
            BlockState &bs = *bsp;
            const std::function<void()> &body = *bodyp;
            Ctx &c = ctx;
            c.tid = uint3_emu{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            c.bid = bid;
            c.bdim = block;
            c.gdim = grid;
            c.block = &bs;
            c.warp = bs.warps[t / 32].get();
            c.lane = t % 32;
            body();
            c.block = nullptr;
            c.warp = nullptr;
        });
    }
}

static std::unique_ptr<BlockState> make_block(dim3 block, size_t smem) {
    const int nthreads = (int)(block.x * block.y * block.z);
    auto bs = std::make_unique<BlockState>(nthreads);
    for (int w = 0; w < (nthreads + 31) / 32; ++w)
        bs->warps.emplace_back(std::make_unique<WarpState>(std::min(32, nthreads - 32 * w)));
    bs->dyn_smem.assign(smem + 16, 0);
    return bs;
}

void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const Mode mode = mode_for(name);
    if (mode == SEQUENTIAL) {
        Ctx &c = ctx;
        c.bdim = block;
        c.gdim = grid;
        c.block = nullptr;
        c.warp = nullptr;
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    c.bid = uint3_emu{bx, by, bz};
                    for (unsigned tz = 0; tz < block.z; ++tz)
                        for (unsigned ty = 0; ty < block.y; ++ty)
                            for (unsigned tx = 0; tx < block.x; ++tx) {
                                c.tid = uint3_emu{tx, ty, tz};
                                c.lane = (int)(tx % 32);
                                body();
                            }
                }
        return;
    }
    if (mode == BLOCK_THREADS) {
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    auto bs = make_block(block, smem);
                    std::vector<std::thread> pool;
                    run_block_threads(*bs, grid, block, uint3_emu{bx, by, bz}, body, pool);
                    for (auto &t : pool) t.join();
                }
        return;
    }
    // CONCURRENT_GRID: every CTA of the launch is alive at once
    std::vector<std::unique_ptr<BlockState>> blocks;
    std::vector<std::thread> pool;
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        blocks.emplace_back(make_block(block, smem));
        run_block_threads(*blocks.back(), grid, block, uint3_emu{bx, 0, 0}, body, pool);
    }
    for (auto &t : pool) t.join();
}
}  // namespace eq_emu

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

cudaError_t cudaDeviceGetAttribute(int *v, int, int) {
    const char *e = getenv("EQ_EMU_SMS");   // "SM count" = concurrent wavefront warps in the emulator
    *v = e ? std::max(1, atoi(e)) : 3;
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new eq_emu_event{0.0};
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = now_ms();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(b->t - a->t);
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
