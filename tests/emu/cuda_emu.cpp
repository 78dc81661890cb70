// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE (see cuda_emu.h).
#include "cuda_emu.h"

#include <chrono>
#include <map>
#include <mutex>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace eq_emu {
thread_local Ctx ctx;

Mode mode_for(const char *name) {
    if (strstr(name, "k_linsolve_exact") || strstr(name, "k_linsolve_tb") || strstr(name, "k_linsolve_wf") || strstr(name, "k_halo_exchange")) return CONCURRENT_GRID;
    if (strstr(name, "k_advect") || strstr(name, "k_divergence_sq") || strstr(name, "k_rb_tiled") || strstr(name, "k_rb_reg") || strstr(name, "k_rb_slide") || strstr(name, "k_rb_stream") || strstr(name, "k_rb_small")) return BLOCK_THREADS;
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
        // one CTA at a time, its threads real OS threads.  The threads are created once per launch and walk the CTAs
        // together: after a CTA every thread arrives at `next_block`, whose completion step installs a fresh
        // BlockState (zeroed shared memory, new barriers) for the following one.
        const int nthreads = (int)(block.x * block.y * block.z);
        const unsigned nblocks = grid.x * grid.y * grid.z;
        if (nblocks == 0 || nthreads == 0) return;
        std::unique_ptr<BlockState> bs = make_block(block, smem);
        auto renew = [&]() noexcept { bs = make_block(block, smem); };
        std::barrier<decltype(renew)> next_block(nthreads, renew);
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
            pool.emplace_back([&, t]() {
                Ctx &c = ctx;
                c.tid = uint3_emu{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                c.bdim = block;
                c.gdim = grid;
                c.lane = t % 32;
                for (unsigned b = 0; b < nblocks; ++b) {
                    c.bid = uint3_emu{b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
                    c.block = bs.get();
                    c.warp = bs->warps[t / 32].get();
                    body();
                    next_block.arrive_and_wait();
                }
                c.block = nullptr;
                c.warp = nullptr;
            });
        for (auto &t : pool) t.join();
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

// ---- shared-memory backed "device" allocations + cudaIpc stand-ins ---------------------------
namespace {
struct ShmBlock {
    std::string name;
    size_t bytes;
    bool owner;
};
std::mutex g_shm_mu;
std::map<void *, ShmBlock> g_shm;
unsigned g_shm_counter = 0;

void shm_cleanup() {
    std::lock_guard<std::mutex> lk(g_shm_mu);
    for (auto &kv : g_shm)
        if (kv.second.owner) shm_unlink(kv.second.name.c_str());
}
}  // namespace

void *eq_emu_shm_alloc(size_t bytes) {
    static bool registered = (atexit(shm_cleanup), true);
    (void)registered;
    const size_t len = (std::max<size_t>(bytes, 1) + 4095) / 4096 * 4096;
    std::lock_guard<std::mutex> lk(g_shm_mu);
    char name[48];
    snprintf(name, sizeof(name), "/eqemu_%d_%u", (int)getpid(), g_shm_counter++);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return nullptr;
    if (ftruncate(fd, (off_t)len) != 0) {
        close(fd);
        shm_unlink(name);
        return nullptr;
    }
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        shm_unlink(name);
        return nullptr;
    }
    g_shm[p] = ShmBlock{name, len, true};
    return p;
}

void eq_emu_shm_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_shm_mu);
    auto it = g_shm.find(p);
    if (it == g_shm.end()) return;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    g_shm.erase(it);
}

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    std::lock_guard<std::mutex> lk(g_shm_mu);
    auto it = g_shm.find(p);
    if (it == g_shm.end()) return 1;
    memset(h, 0, sizeof(*h));
    snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    const uint64_t bytes = it->second.bytes;
    memcpy(h->reserved + 48, &bytes, 8);
    return cudaSuccess;
}

cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    uint64_t bytes = 0;
    memcpy(&bytes, h.reserved + 48, 8);
    const int fd = shm_open(h.reserved, O_RDWR, 0600);
    if (fd < 0) return 1;
    void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return 1;
    std::lock_guard<std::mutex> lk(g_shm_mu);
    g_shm[m] = ShmBlock{h.reserved, (size_t)bytes, false};
    *p = m;
    return cudaSuccess;
}

cudaError_t cudaIpcCloseMemHandle(void *p) {
    eq_emu_shm_free(p);
    return cudaSuccess;
}
