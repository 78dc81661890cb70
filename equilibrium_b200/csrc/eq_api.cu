// eq_api.cu -- C ABI of libequilibrium_cuda.so (include/equilibrium_cuda.h):
// handle, device memory, stream-ordered composition of Fluid::step
// (fluid.rs:437-524) out of the kernels in k_*.cuh.
#include "../../include/equilibrium_cuda.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <cmath>
#include <new>
#include <vector>

#include "eq_common.cuh"
#include "k_linsolve_exact.cuh"
#include "k_linsolve_tb.cuh"
#include "k_linsolve_wf.cuh"
#include "k_linsolve_rb.cuh"
#include "k_linsolve_rbs.cuh"
#include "k_stencils.cuh"
#include "k_multigpu.cuh"
#include "k_linsolve_rbsmall.cuh"

#include <unistd.h>

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Knobs that change RESULTS (dependency waits off, stage-isolation builds) exist only in builds made with
// -DEQ_DEBUG_KNOBS (scripts/); the default library ignores them.
static int debug_knob(const char *name) {
#ifdef EQ_DEBUG_KNOBS
    return getenv(name) ? 1 : 0;
#else
    (void)name;
    return 0;
#endif
}

static thread_local char g_err[512] = "";

static int eq_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return eq_fail(EQ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                           __FILE__, __LINE__);                                                    \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int r_ = (expr);          \
        if (r_ != EQ_OK) return r_; \
    } while (0)

#define NEED(h)                                                      \
    do {                                                             \
        if (!(h)) return eq_fail(EQ_ERR_INVALID, "null handle");     \
        CU(cudaSetDevice((h)->dev));                                 \
    } while (0)

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
enum { CAT_LS = 0, CAT_ADV, CAT_PROJ, CAT_BND, CAT_OTHER, CAT_COUNT };

struct ProfSpan {
    int cat;
    cudaEvent_t a, b;
};

#define LSX_KMAX 256   // iterations per wavefront launch (bounds the job table and flag array)

struct eq_fluid {
    EqParams prm;
    EqLayout L;
    int dev;
    int sm_count;
    cudaStream_t own_stream, stream;
    float *f[6];            // EQ_F_DENSITY .. EQ_F_SCRATCH
    float *rb_tmp;          // ping-pong partner of x in the tiled red-black solver
    uint8_t *cells;
    // mask-derived tables (rebuilt lazily when the mask changed)
    uint8_t *codes, *row_fluid, *col_fluid, *chunk_flags, *chunk_flags_tb;
    float *tb_raw[2], *tb_edge[2];   // side streams of the temporally blocked solver
    int tb_ctas;
    uint8_t *wf_flags;               // register wavefront solver (k_linsolve_wf.cuh): (orientation, band, chunk) summaries,
    float *wf_raw[2], *wf_edge[2];   // its side streams
    int wf_ctas;
    unsigned *a0_flags;              // [2] guard results of the a == 0 shortcut (k_a0_check), one per request
    const unsigned *run_if;          // set around the solver launches of a request whose shortcut is armed
    unsigned long long *wf_trace;    // EQ_WF_TRACE=1
    unsigned *wf_dbg;                // EQ_WF_DEBUG=1
    int wf_dbg_ctas;
    unsigned *counts;       // [4] device
    bool all_cols_fluid;
    uint2 *row_list, *col_list;
    unsigned n_row, n_col;
    size_t cap_row, cap_col;
    bool mask_dirty;
    // exact wavefront workspace
    float *raw[2];
    unsigned *flags;        // [0] ticket, [1] error, [8..] progress
    size_t flags_words;
    std::map<int, uint32_t *> *job_tables;   // keyed by iterations-per-launch
    int lsx_ctas;
    unsigned long long *lsx_trace;
    unsigned long long *lsx_jobtimes;
    size_t lsx_jobtimes_n;
    unsigned long long *lsx_stats;   // EQ_LSX_STATS=1: device cycle counters of the wavefront kernel
    // timing / profiling
    cudaEvent_t ev0, ev1;
    bool prof_on;
    std::vector<ProfSpan> *spans;
    std::vector<cudaEvent_t> *ev_pool;
    double prof_ms[CAT_COUNT];
    int64_t prof_launches[CAT_COUNT];
    int64_t prof_cell_iters;
    int64_t prof_steps;
    void *l2buf;
    size_t l2bytes;
    // row slabs over several GPUs (one process per GPU, or several handles in one process)
    int rank, world;
    int b_lo, b_hi;                              // owned bands of the wavefront solver
    float *peer_f[EQ_MAX_RANKS][6];              // every rank's fields / raw streams / flags / sync slots
    float *peer_raw[EQ_MAX_RANKS][2];
    float *peer_tmp[EQ_MAX_RANKS];
    unsigned *peer_flags[EQ_MAX_RANKS];
    unsigned *peer_sync[EQ_MAX_RANKS];
    bool peer_ipc[EQ_MAX_RANKS];                 // mapped with cudaIpcOpenMemHandle (must be closed)
    unsigned *sync;                              // my cross-GPU sync slots
    unsigned halo_epoch, bar_epoch, rb_epoch, or_epoch;
    bool attached;
    // per-frame snapshots: staging slots, their events, the copy stream
    cudaStream_t copy_stream;
    void *snap_buf[EQ_SNAPSHOT_SLOTS];
    cudaEvent_t snap_ready[EQ_SNAPSHOT_SLOTS], snap_done[EQ_SNAPSHOT_SLOTS];
    bool snap_used[EQ_SNAPSHOT_SLOTS];
};

static size_t field_elems(const eq_fluid *h) { return (size_t)h->L.P * h->L.rows; }

struct ProfScope {
    eq_fluid *h;
    int cat;
    bool timed;
    cudaEvent_t a, b;
    ProfScope(eq_fluid *h_, int cat_, int launches) : h(h_), cat(cat_), timed(false) {
        h->prof_launches[cat] += launches;
        if (h->prof_on) {
            auto get = [&]() {
                cudaEvent_t e;
                if (!h->ev_pool->empty()) {
                    e = h->ev_pool->back();
                    h->ev_pool->pop_back();
                } else {
                    cudaEventCreate(&e);
                }
                return e;
            };
            a = get();
            b = get();
            cudaEventRecord(a, h->stream);
            timed = true;
        }
    }
    ~ProfScope() {
        if (timed) {
            cudaEventRecord(b, h->stream);
            h->spans->push_back(ProfSpan{cat, a, b});
        }
    }
};

static int prof_collect(eq_fluid *h) {
    if (h->spans->empty()) return EQ_OK;
    CU(cudaStreamSynchronize(h->stream));
    for (auto &s : *h->spans) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.a, s.b));
        h->prof_ms[s.cat] += ms;
        h->ev_pool->push_back(s.a);
        h->ev_pool->push_back(s.b);
    }
    h->spans->clear();
    return EQ_OK;
}

static inline dim3 row_grid(const eq_fluid *h, int rows, int threads = 256) {
    return dim3((unsigned)((h->L.N + threads - 1) / threads), (unsigned)rows, 1);
}
// k_divergence / k_gradient: a block covers 4 * EQ_ST_THREADS columns x EQ_ST_ROWS rows
static inline dim3 stencil_grid(const eq_fluid *h, int rows) {
    const int cols_per_block = 4 * EQ_ST_THREADS;
    return dim3((unsigned)((h->L.N + cols_per_block - 1) / cols_per_block), (unsigned)((rows + EQ_ST_ROWS - 1) / EQ_ST_ROWS), 1);
}
static int check_launch(const char *what);
// rows of the interior (1..N-2) that this rank owns
static inline int owned_interior_rows(const eq_fluid *h) {
    return std::min(h->L.row1, h->L.N - 1) - std::max(h->L.row0, 1);
}

// ---------------------------------------------------------------------------
// multi-GPU plumbing: halo rows and barriers (k_multigpu.cuh)
// ---------------------------------------------------------------------------
static int field_index(const eq_fluid *h, const float *f) {
    for (int i = 0; i < 6; ++i)
        if (h->f[i] == f) return i;
    return -1;
}

static int need_attached(eq_fluid *h) {
    if (h->world > 1 && !h->attached)
        return eq_fail(EQ_ERR_STATE, "multi-GPU handle used before eq_ipc_attach (rank %d of %d)", h->rank, h->world);
    return EQ_OK;
}

// refresh the ghost rows of a buffer on both neighbours and mine from theirs
static int halo_xchg_buf(eq_fluid *h, float *buf, float *up, float *down, int nrows) {
    ProfScope ps(h, CAT_OTHER, 1);
    EqHaloArgs a;
    memset(&a, 0, sizeof(a));
    a.field = buf;
    a.peer_up = up;
    a.peer_down = down;
    a.sync = h->sync;
    a.sync_up = h->rank > 0 ? h->peer_sync[h->rank - 1] : nullptr;
    a.sync_down = h->rank + 1 < h->world ? h->peer_sync[h->rank + 1] : nullptr;
    a.epoch = ++h->halo_epoch;
    a.nrows = nrows;
    a.error = reinterpret_cast<int *>(h->flags + 1);
    a.run_if = h->run_if;
    EQ_LAUNCH(k_halo_exchange, 2, 1024, 16, h->stream, a, h->L);
    return check_launch("k_halo_exchange");
}

static int halo_xchg(eq_fluid *h, float *field, int nrows = 1) {
    if (h->world <= 1) return EQ_OK;
    TRY(need_attached(h));
    const bool up = h->rank > 0, down = h->rank + 1 < h->world;
    if (field == h->rb_tmp)
        return halo_xchg_buf(h, field, up ? h->peer_tmp[h->rank - 1] : nullptr, down ? h->peer_tmp[h->rank + 1] : nullptr, nrows);
    const int fi = field_index(h, field);
    if (fi < 0) return eq_fail(EQ_ERR_INVALID, "halo exchange of an unknown field");
    return halo_xchg_buf(h, field, up ? h->peer_f[h->rank - 1][fi] : nullptr, down ? h->peer_f[h->rank + 1][fi] : nullptr, nrows);
}

// every rank has finished what it launched so far
static int barrier_all(eq_fluid *h) {
    if (h->world <= 1) return EQ_OK;
    TRY(need_attached(h));
    ProfScope ps(h, CAT_OTHER, 1);
    EqBarrierArgs a;
    memset(&a, 0, sizeof(a));
    a.sync = h->sync;
    for (int r = 0; r < h->world; ++r) a.peer_sync[r] = h->peer_sync[r];
    a.rank = h->rank;
    a.world = h->world;
    a.epoch = ++h->bar_epoch;
    a.error = reinterpret_cast<int *>(h->flags + 1);
    EQ_LAUNCH(k_barrier_all, 1, 32, 0, h->stream, a);
    return check_launch("k_barrier_all");
}

static EqPeerTable peer_table(const eq_fluid *h, const float *field) {
    EqPeerTable t;
    memset(&t, 0, sizeof(t));
    t.world = std::max(1, h->world);
    const int fi = field ? field_index(h, field) : -1;
    const int NB = (h->L.N - 2 + 31) / 32;
    for (int r = 0; r < t.world; ++r) {
        t.base[r] = (h->world > 1 && fi >= 0) ? h->peer_f[r][fi] : field;
        t.row_begin[r] = (r == 0) ? 0 : 1 + 32 * (int)((int64_t)r * NB / t.world);
    }
    t.row_begin[t.world] = h->L.N;
    return t;
}

static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return eq_fail(EQ_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return EQ_OK;
}

// ---------------------------------------------------------------------------
// mask tables
// ---------------------------------------------------------------------------
static int ensure_tables(eq_fluid *h) {
    if (!h->mask_dirty) return EQ_OK;
    ProfScope ps(h, CAT_OTHER, 2);
    const EqLayout L = h->L;
    CU(cudaMemsetAsync(h->counts, 0, 4 * sizeof(unsigned), h->stream));
    CU(cudaMemsetAsync(h->row_fluid, 0, L.N, h->stream));
    CU(cudaMemsetAsync(h->col_fluid, 0, L.P, h->stream));
    CU(cudaMemsetAsync(h->chunk_flags, 0, 2 * (size_t)((L.N - 2 + 31) / 32) * ((L.N + EQ_LSX_CW - 1) / EQ_LSX_CW), h->stream));
    CU(cudaMemsetAsync(h->chunk_flags_tb, 0, 2 * (size_t)((L.N - 2 + TBX_SK + 31) / 32) * ((L.N + EQ_LSX_CW - 1) / EQ_LSX_CW), h->stream));
    EQ_LAUNCH(k_build_codes, row_grid(h, L.N), 256, 0, h->stream, h->cells, h->codes, h->row_fluid, h->col_fluid,
              h->chunk_flags, h->chunk_flags_tb, h->counts, nullptr, nullptr, 0, L);
    TRY(check_launch("k_build_codes"));
    unsigned counts[4];
    CU(cudaMemcpyAsync(counts, h->counts, sizeof(counts), cudaMemcpyDeviceToHost, h->stream));
    std::vector<uint8_t> colf((size_t)L.N);
    CU(cudaMemcpyAsync(colf.data(), h->col_fluid, (size_t)L.N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // every interior column holds a NoWall cell (anything but an obstacle spanning the full height): the Passive
    // frame-row copies of quirk Q6 are then unconditional and the wavefront solver keeps its fast loop for them
    h->all_cols_fluid = true;
    for (int i = 1; i <= L.N - 2; ++i) h->all_cols_fluid = h->all_cols_fluid && colf[(size_t)i] != 0;
    if (counts[0] > h->cap_row) {
        if (h->row_list) CU(cudaFree(h->row_list));
        h->row_list = nullptr;
        h->cap_row = (size_t)counts[0] + counts[0] / 4 + 1024;
        CU(cudaMalloc(&h->row_list, h->cap_row * sizeof(uint2)));
    }
    if (counts[1] > h->cap_col) {
        if (h->col_list) CU(cudaFree(h->col_list));
        h->col_list = nullptr;
        h->cap_col = (size_t)counts[1] + counts[1] / 4 + 1024;
        CU(cudaMalloc(&h->col_list, h->cap_col * sizeof(uint2)));
    }
    EQ_LAUNCH(k_build_codes, row_grid(h, L.N), 256, 0, h->stream, h->cells, h->codes, h->row_fluid, h->col_fluid,
              h->chunk_flags, h->chunk_flags_tb, h->counts, h->row_list, h->col_list, 1, L);
    TRY(check_launch("k_build_codes(lists)"));
    {
        const int NBPw = (L.N + WF_SK + 31) / 32, NCw = L.P / WF_CW;
        CU(cudaMemsetAsync(h->wf_flags, 0, 3 * (size_t)NBPw * NCw, h->stream));
        EQ_LAUNCH(k_build_wf_flags, row_grid(h, L.N), 256, 0, h->stream, h->codes, h->wf_flags, NBPw, NCw, L);
        TRY(check_launch("k_build_wf_flags"));
    }
    h->n_row = counts[0];
    h->n_col = counts[1];
    h->mask_dirty = false;
    return EQ_OK;
}

// ---------------------------------------------------------------------------
// set_boundaries (fluid.rs:252-272) as a sparse pass
// ---------------------------------------------------------------------------
static int set_boundaries(eq_fluid *h, int orient, float *x) {
    TRY(ensure_tables(h));
    if (orient == EQ_ADJUST_COLUMN) TRY(halo_xchg(h, x));   // the mirrored wall cell may sit in a ghost row
    ProfScope ps(h, CAT_BND, 1);
    const EqLayout L = h->L;
    if (orient == EQ_PASSIVE) {
        const int threads = 256, blocks = (L.N + threads - 1) / threads;
        EQ_LAUNCH(k_bnd_passive, blocks, threads, 0, h->stream, x, h->row_fluid, h->col_fluid, L);
        return check_launch("k_bnd_passive");
    }
    const uint2 *list = orient == EQ_ADJUST_ROW ? h->row_list : h->col_list;
    const unsigned n = orient == EQ_ADJUST_ROW ? h->n_row : h->n_col;
    const int threads = 256;
    const unsigned blocks = std::max(1u, (n + threads - 1) / threads);
    EQ_LAUNCH(k_bnd_list, blocks, threads, 0, h->stream, x, list, n, L);
    return check_launch("k_bnd_list");
}

// ---------------------------------------------------------------------------
// lin_solve (fluid.rs:301-325)
// ---------------------------------------------------------------------------
struct LinSolveReq {
    int orient;
    float *x;
    const float *x0;
    float a, c;
};

// Iterations are handed out in groups of `kgroup`: inside a group jobs go in wavefront order
// w = b + 2(k - k0) (a job depends only on w-1, or on the previous group).  One group = every
// iteration is the plain wavefront; small groups let the sweep reach the last band of a slab
// early, which is what lets the next GPU start (pipeline over groups).
static int job_group_size(const eq_fluid *h, int kc) {
    if (const char *e = getenv("EQ_LSX_KGROUP")) return std::max(1, std::min(kc, atoi(e)));
    if (h->world <= 1) return kc;
    return std::max(1, std::min(kc, (kc + 2 * h->world - 1) / (2 * h->world) + 1));
}

static int get_job_table(eq_fluid *h, int kc, const uint32_t **out) {
    const int G = job_group_size(h, kc);
    const int key = kc * 1024 + G;
    auto it = h->job_tables->find(key);
    if (it != h->job_tables->end()) {
        *out = it->second;
        return EQ_OK;
    }
    const int NB = (h->L.N - 2 + 31) / 32;
    std::vector<uint32_t> tab;
    tab.reserve((size_t)kc * (h->b_hi - h->b_lo));
    // A rank lists only the bands of its slab; every rank uses the same global order, so cross-GPU
    // waits also point at jobs that were handed out earlier on their own GPU.
    for (int k0 = 0; k0 < kc; k0 += G) {
        const int k1 = std::min(kc, k0 + G);
        for (int w = 0; w <= (NB - 1) + 2 * (k1 - k0 - 1); ++w)
            for (int k = k0; k < k1; ++k) {
                const int b = w - 2 * (k - k0);
                if (b >= h->b_lo && b < h->b_hi) tab.push_back(((uint32_t)k << 16) | (uint32_t)b);
            }
    }
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, tab.size() * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // `tab` is pageable and goes out of scope
    (*h->job_tables)[key] = d;
    *out = d;
    return EQ_OK;
}

static int lin_solve_exact(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    const EqLayout L = h->L;
    const int NB = (L.N - 2 + 31) / 32;
    const int NC = (L.N + EQ_LSX_CW - 1) / EQ_LSX_CW;
    int64_t done = 0;
    while (done < iters) {
        const int kc = (int)std::min<int64_t>(LSX_KMAX, iters - done);
        const uint32_t *jobs = nullptr;
        TRY(get_job_table(h, kc, &jobs));
        LsxParams p;
        memset(&p, 0, sizeof(p));
        p.nprob = nreq;
        const size_t prog_words = (size_t)kc * NB;
        for (int i = 0; i < nreq; ++i) {
            p.prob[i].x = req[i].x;
            p.prob[i].x0 = req[i].x0;
            p.prob[i].raw = h->raw[i];
            p.prob[i].progress = h->flags + 8 + (size_t)i * prog_words;
            p.prob[i].a = req[i].a;
            p.prob[i].c_recip = 1.0f / req[i].c;                       // fluid.rs:311
            p.prob[i].orient = req[i].orient;
            if (h->world > 1) {
                const int fi = field_index(h, req[i].x);
                if (fi < 0) return eq_fail(EQ_ERR_INVALID, "multi-GPU lin_solve needs one of the handle's fields");
                if (h->rank > 0) {
                    p.prob[i].x_up = h->peer_f[h->rank - 1][fi];
                    p.prob[i].prog_up = h->peer_flags[h->rank - 1] + 8 + (size_t)i * prog_words;
                }
                if (h->rank + 1 < h->world) {
                    p.prob[i].raw_down = h->peer_raw[h->rank + 1][i];
                    p.prob[i].prog_down = h->peer_flags[h->rank + 1] + 8 + (size_t)i * prog_words;
                }
            }
        }
        p.b_lo = h->b_lo;
        p.b_hi = h->b_hi;
        p.codes = h->codes;
        p.chunk_flags = h->chunk_flags;
        p.row_fluid = h->row_fluid;
        p.col_fluid = h->col_fluid;
        p.jobs = jobs;
        p.njobs = kc * (h->b_hi - h->b_lo);
        p.N = L.N;
        p.P = L.P;
        p.K = kc;
        p.NB = NB;
        p.NC = NC;
        p.ticket = h->flags;
        p.error = reinterpret_cast<int *>(h->flags + 1);
        p.stats = h->lsx_stats;
        p.debug_nodeps = debug_knob("EQ_LSX_NODEPS");
        p.rotate_roles = env_int("EQ_LSX_ROT", 1);
        p.pub_batch = std::max(1, env_int("EQ_LSX_PUBBATCH", L.N >= 8192 ? 8 : 4));
        p.slack = getenv("EQ_LSX_SLACK") ? atoi(getenv("EQ_LSX_SLACK")) : 0;
        p.trace = h->lsx_trace;
        p.jobtimes = nullptr;
        if (getenv("EQ_LSX_JOBTIMES")) {
            const size_t n = 2 * (size_t)kc * NB * nreq;
            if (h->lsx_jobtimes_n < n) {
                cudaFree(h->lsx_jobtimes);
                CU(cudaMalloc(&h->lsx_jobtimes, n * sizeof(unsigned long long)));
                h->lsx_jobtimes_n = n;
            }
            CU(cudaMemsetAsync(h->lsx_jobtimes, 0, n * sizeof(unsigned long long), h->stream));
            p.jobtimes = h->lsx_jobtimes;
        }
        // ticket := 0, progress := 0; the sticky error word is left alone
        CU(cudaMemsetAsync(h->flags, 0, sizeof(unsigned), h->stream));
        CU(cudaMemsetAsync(h->flags + 8, 0, (size_t)nreq * prog_words * sizeof(unsigned), h->stream));
        // Several GPUs: the exchange (a) gives my last band the neighbour's first row as its initial
        // F_{-1} and (b) is a barrier with both neighbours, so nobody publishes into a progress
        // mirror that has not been cleared yet.
        for (int i = 0; i < nreq; ++i) TRY(halo_xchg(h, req[i].x));
        // odd number of phases per mbarrier slot and job (NC not a multiple of 2 * slots): one job per CTA
        p.single_shot = (NC % (2 * LSX_SLOTS)) != 0 ? 1 : 0;
        p.run_if = h->run_if;
        const int grid = p.single_shot ? p.njobs * nreq : std::min(h->lsx_ctas, p.njobs * nreq);
        EQ_LAUNCH(k_linsolve_exact, grid, LSX_THREADS, LSX_SMEM_BYTES, h->stream, p);
        TRY(check_launch("k_linsolve_exact"));
        // ... and afterwards it waits until the neighbour's solver (which patches my last row and
        // publishes into my mirrors until its very end) is done, then refreshes the ghost rows.
        for (int i = 0; i < nreq; ++i) TRY(halo_xchg(h, req[i].x));
        done += kc;
    }
    for (int i = 0; i < nreq; ++i) {
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
    }
    return EQ_OK;
}

// Temporally blocked variant (k_linsolve_tb.cuh): TBX_T iterations per job.  Single GPU.
static int get_tb_job_table(eq_fluid *h, int G, const uint32_t **out) {
    const int key = -G;   // shares the cache with the plain tables (their keys are positive)
    auto it = h->job_tables->find(key);
    if (it != h->job_tables->end()) {
        *out = it->second;
        return EQ_OK;
    }
    const int NBP = (h->L.N - 2 + TBX_SK + 31) / 32;
    std::vector<uint32_t> tab;
    tab.reserve((size_t)G * NBP);
    for (int w = 0; w <= (NBP - 1) + 2 * (G - 1); ++w)      // w = b + 2g: both dependencies have w-1
        for (int g = 0; g < G; ++g) {
            const int b = w - 2 * g;
            if (b >= 0 && b < NBP) tab.push_back(((uint32_t)g << 16) | (uint32_t)b);
        }
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, tab.size() * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    (*h->job_tables)[key] = d;
    *out = d;
    return EQ_OK;
}

static int lin_solve_exact_tb(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    // Two independent solves (the velocity diffusions) can share one launch, but measured on 16384^2 K=20 the
    // shared launch takes 53.9 ms against 16.6 + 19.0 ms back to back: each problem gets half of the resident
    // CTAs while its band-to-band chain stays as long, and two orientations' loops compete for the
    // instruction cache.  One launch per field unless EQ_TB_BATCH=1.
    if (nreq > 1 && !env_int("EQ_TB_BATCH", 0)) {
        for (int i = 0; i < nreq; ++i) TRY(lin_solve_exact_tb(h, req + i, 1, iters));
        return EQ_OK;
    }
    const EqLayout L = h->L;
    const int NBP = (L.N - 2 + TBX_SK + 31) / 32;
    const int NC = (L.N + EQ_LSX_CW - 1) / EQ_LSX_CW;
    int64_t done = 0;
    while (done < iters) {
        const int kc = (int)std::min<int64_t>(LSX_KMAX, iters - done);
        const int G = (kc + TBX_T - 1) / TBX_T;
        const uint32_t *jobs = nullptr;
        TRY(get_tb_job_table(h, G, &jobs));
        TbxParams p;
        memset(&p, 0, sizeof(p));
        p.nprob = nreq;
        const size_t prog_words = (size_t)G * NBP;
        for (int i = 0; i < nreq; ++i) {
            p.prob[i].x = req[i].x;
            p.prob[i].x0 = req[i].x0;
            p.prob[i].raw = h->tb_raw[i];
            p.prob[i].edge = h->tb_edge[i];
            p.prob[i].progress = h->flags + 8 + (size_t)i * prog_words;
            p.prob[i].a = req[i].a;
            p.prob[i].c_recip = 1.0f / req[i].c;                       // fluid.rs:311
            p.prob[i].orient = req[i].orient;
        }
        p.codes = h->codes;
        p.chunk_flags = h->chunk_flags_tb;
        p.row_fluid = h->row_fluid;
        p.col_fluid = h->col_fluid;
        p.jobs = jobs;
        p.njobs = G * NBP;
        p.N = L.N;
        p.P = L.P;
        p.K = kc;
        p.G = G;
        p.NBP = NBP;
        p.NC = NC;
        p.ticket = h->flags;
        p.error = reinterpret_cast<int *>(h->flags + 1);
        p.rotate_roles = env_int("EQ_LSX_ROT", 1);
        p.pub_batch = std::max(1, env_int("EQ_LSX_PUBBATCH", L.N >= 8192 ? 4 : 2));
        p.debug_nodeps = debug_knob("EQ_LSX_NODEPS");
        p.passive_fast_frames = (h->all_cols_fluid && env_int("EQ_TB_FAST_FRAMES", 1)) ? 1 : 0;
        p.trace = h->lsx_trace;
        p.trace_g = G - 1;
        p.trace_b = 200;
        p.trace_q = 200;
        if (const char *e = getenv("EQ_LSX_TRACE")) {
            int tg = 0, tb = 0, tq = 200;
            if (sscanf(e, "%d,%d,%d", &tg, &tb, &tq) >= 2) { p.trace_g = std::min(tg, G - 1); p.trace_b = tb; p.trace_q = tq; }
        }
        p.jobtimes = nullptr;
        if (getenv("EQ_LSX_JOBTIMES")) {
            const size_t n = 4 * (size_t)G * NBP;
            if (h->lsx_jobtimes_n < n) {
                cudaFree(h->lsx_jobtimes);
                CU(cudaMalloc(&h->lsx_jobtimes, n * sizeof(unsigned long long)));
                h->lsx_jobtimes_n = n;
            }
            CU(cudaMemsetAsync(h->lsx_jobtimes, 0, n * sizeof(unsigned long long), h->stream));
            p.jobtimes = h->lsx_jobtimes;
        }
        CU(cudaMemsetAsync(h->flags, 0, sizeof(unsigned), h->stream));
        CU(cudaMemsetAsync(h->flags + 8, 0, (size_t)nreq * prog_words * sizeof(unsigned), h->stream));
        p.single_shot = (NC % (2 * TBX_SLOTS)) != 0 ? 1 : 0;   // odd number of phases per mbarrier slot and job: one job per CTA
        p.run_if = h->run_if;
        const int grid = p.single_shot ? p.njobs * nreq : std::min(h->tb_ctas, p.njobs * nreq);
        EQ_LAUNCH(k_linsolve_tb, grid, TBX_THREADS, TBX_SMEM_BYTES, h->stream, p);
        TRY(check_launch("k_linsolve_tb"));
        done += kc;
    }
    for (int i = 0; i < nreq; ++i) {
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
    }
    return EQ_OK;
}

// Exact mode on row slabs, plan "replica": the band-to-band chain of the wavefront solver does not get shorter when
// the rows are spread over GPUs, and the row-slab kernel is the older one without fused iterations: with 2
// GPUs a slab solve is much slower than the single-GPU kernel on the whole grid.  So every rank gathers the other
// slabs of x and x0 over NVLink (direct peer loads, ~2.5 ms per solve at 16384^2 on 2 GPUs) and runs k_linsolve_tb on the
// whole grid; all ranks end up with the same, complete x, and the stencils around the solve stay slab-parallel.
// Measured C4 frame: 77.8 ms on 2 GPUs (slabs 96.8; one GPU 72.1), 88.5 ms on 4 (slabs 68): the default for 2 ranks only.
static int lin_solve_exact_replica(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    TRY(need_attached(h));
    const EqLayout L = h->L;
    {
        ProfScope ps(h, CAT_OTHER, 2 + 2 * nreq * (h->world - 1));
        TRY(barrier_all(h));                                    // every slab of x and x0 is final
        const EqPeerTable t = peer_table(h, req[0].x);          // (row partition)
        for (int i = 0; i < nreq; ++i) {
            const float *src[2] = {req[i].x, req[i].x0};
            for (int q = 0; q < h->world; ++q) {
                if (q == h->rank) continue;
                const int r0 = t.row_begin[q], r1 = t.row_begin[q + 1];
                for (int f = 0; f < 2; ++f) {
                    const int fi = field_index(h, src[f]);
                    if (fi < 0) return eq_fail(EQ_ERR_INVALID, "replicated solve of an unknown field");
                    const dim3 g((unsigned)((L.P / 4 + 255) / 256), (unsigned)std::min(r1 - r0, 1024), 1);
                    EQ_LAUNCH(k_copy_rows_if, g, 256, 0, h->stream, h->f[fi], h->peer_f[q][fi], h->run_if, r0, r1, L);
                    TRY(check_launch("k_copy_rows_if"));
                }
            }
        }
        TRY(barrier_all(h));                                    // nobody overwrites rows that a peer is still reading
    }
    const EqLayout mine = h->L;
    h->L.row0 = 0;
    h->L.row1 = L.N;
    const int rc = lin_solve_exact_tb(h, req, nreq, iters);
    h->L = mine;
    return rc;
}

#ifndef RQ_SEG_ROWS
#define RQ_SEG_ROWS 192           // rows per k_rb_stream task (see lin_solve_red_black_stream)
#endif
#ifndef EQ_EXACT_REPLICA_MAX_WORLD
#define EQ_EXACT_REPLICA_MAX_WORLD 2   // C4 frame on 2 / 4 / 8 GPUs: slabs 96.8 / 68 / 49.6 ms, replicated solve 77.8 / 88.5 / - (the gather grows with the ranks)
#endif
#ifndef EQ_RB_STREAM_MIN_N
#define EQ_RB_STREAM_MIN_N 2048   // C3 (4096^2): k_rb_stream 1.47 ms / solve, k_rb_slide 2.14, k_rb_reg 2.30; C2 (1024^2): k_rb_reg 0.32 ms
#endif
#ifndef EQ_DEFAULT_EXACT_WF
#define EQ_DEFAULT_EXACT_WF 0   // until k_linsolve_wf beats k_linsolve_tb on the BASELINE configs
#endif
// Register-blocked wavefront (k_linsolve_wf.cuh): WF_T iterations per job, passed between the iterations of a job in
// registers.  Single GPU.
static int get_wf_job_table(eq_fluid *h, int G, const uint32_t **out) {
    const int key = -(1 << 20) - G;   // shares the cache with the other tables
    auto it = h->job_tables->find(key);
    if (it != h->job_tables->end()) {
        *out = it->second;
        return EQ_OK;
    }
    const int NBP = (h->L.N + WF_SK + 31) / 32;
    std::vector<uint32_t> tab;
    tab.reserve((size_t)G * NBP);
    for (int w = 0; w <= (NBP - 1) + 2 * (G - 1); ++w)      // w = b + 2g: both dependencies have w-1
        for (int g = 0; g < G; ++g) {
            const int b = w - 2 * g;
            if (b >= 0 && b < NBP) tab.push_back(((uint32_t)g << 16) | (uint32_t)b);
        }
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, tab.size() * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    (*h->job_tables)[key] = d;
    *out = d;
    return EQ_OK;
}

static int lin_solve_exact_wf(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    if (nreq > 1) {   // one launch per field: a shared launch halves the resident jobs of each chain (see lin_solve_exact_tb)
        for (int i = 0; i < nreq; ++i) TRY(lin_solve_exact_wf(h, req + i, 1, iters));
        return EQ_OK;
    }
    const EqLayout L = h->L;
    const int NBP = (L.N + WF_SK + 31) / 32;
    const int NC = L.P / WF_CW;
    const int kmax = LSX_KMAX / WF_T * WF_T;
    int64_t done = 0;
    while (done < iters) {
        const int kc = (int)std::min<int64_t>(kmax, iters - done);
        const int G = (kc + WF_T - 1) / WF_T;
        const uint32_t *jobs = nullptr;
        TRY(get_wf_job_table(h, G, &jobs));
        WfParams p;
        memset(&p, 0, sizeof(p));
        p.nprob = nreq;
        const size_t prog_words = (size_t)G * NBP;
        if (8 + (size_t)nreq * prog_words > h->flags_words) return eq_fail(EQ_ERR_INVALID, "progress table too small");
        for (int i = 0; i < nreq; ++i) {
            p.prob[i].x = req[i].x;
            p.prob[i].x0 = req[i].x0;
            p.prob[i].raw = h->wf_raw[i];
            p.prob[i].edge = h->wf_edge[i];
            p.prob[i].progress = h->flags + 8 + (size_t)i * prog_words;
            p.prob[i].a = req[i].a;
            p.prob[i].c_recip = 1.0f / req[i].c;                       // fluid.rs:311
            p.prob[i].orient = req[i].orient;
        }
        p.codes = h->codes;
        p.row_fluid = h->row_fluid;
        p.flags = h->wf_flags;
        p.jobs = jobs;
        p.njobs = G * NBP;
        p.N = L.N;
        p.P = L.P;
        p.K = kc;
        p.G = G;
        p.NBP = NBP;
        p.NC = NC;
        p.ticket = h->flags;
        p.error = reinterpret_cast<int *>(h->flags + 1);
        p.rotate_roles = env_int("EQ_LSX_ROT", 1);
        p.pub_batch = std::max(1, env_int("EQ_WF_PUBBATCH", L.N >= 8192 ? 4 : 2));
        p.force_general = env_int("EQ_WF_GENERAL", 0);
        p.run_if = h->run_if;
        p.debug_nodeps = debug_knob("EQ_LSX_NODEPS");
        CU(cudaMemsetAsync(h->flags, 0, sizeof(unsigned), h->stream));
        CU(cudaMemsetAsync(h->flags + 8, 0, (size_t)nreq * prog_words * sizeof(unsigned), h->stream));
        const int grid = std::min(h->wf_ctas, p.njobs * nreq);
        unsigned long long *jt = nullptr;
        if (getenv("EQ_WF_JOBTIMES")) {
            CU(cudaMalloc(&jt, 4 * (size_t)p.njobs * sizeof(unsigned long long)));
            CU(cudaMemsetAsync(jt, 0, 4 * (size_t)p.njobs * sizeof(unsigned long long), h->stream));
            p.jobtimes = jt;
        }
        if (getenv("EQ_WF_TRACE")) {
            if (!h->wf_trace) CU(cudaMalloc(&h->wf_trace, 64 * 16 * sizeof(unsigned long long)));
            CU(cudaMemsetAsync(h->wf_trace, 0, 64 * 16 * sizeof(unsigned long long), h->stream));
            p.trace = h->wf_trace;
        }
        if (getenv("EQ_WF_DEBUG")) {
            if (!h->wf_dbg) CU(cudaMalloc(&h->wf_dbg, (size_t)h->wf_ctas * 32 * sizeof(unsigned)));
            CU(cudaMemsetAsync(h->wf_dbg, 0, (size_t)h->wf_ctas * 32 * sizeof(unsigned), h->stream));
            h->wf_dbg_ctas = grid;
            p.dbg = h->wf_dbg;
        }
        EQ_LAUNCH(k_linsolve_wf, grid, WF_THREADS, WF_SMEM_BYTES, h->stream, p);
        TRY(check_launch("k_linsolve_wf"));
        if (jt) {
            std::vector<unsigned long long> t(4 * (size_t)p.njobs);
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaMemcpy(t.data(), jt, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            cudaFree(jt);
            if (FILE *f = fopen(getenv("EQ_WF_JOBTIMES"), "w")) {
                fprintf(f, "# G=%d NBP=%d ; g b ticket_ns flags_ns full0_ns done_ns\n", G, NBP);
                for (int gg = 0; gg < G; ++gg)
                    for (int bb = 0; bb < NBP; ++bb) {
                        const unsigned long long *e = t.data() + 4 * ((size_t)gg * NBP + bb);
                        fprintf(f, "%d %d %llu %llu %llu %llu\n", gg, bb, e[0], e[1], e[2], e[3]);
                    }
                fclose(f);
            }
        }
        if (p.trace) {
            static const char *names[16] = {"loader.start", "flags0.ok", "wave0.issued", "full0.seen", "macro0.done", "macro2.done",
                                            "macro3.done", "chunk0.stored", "chunk0.published", "flags4.ok", "macro7.done",
                                            "macro63.done", "macro127.done", "-", "-", "-"};
            std::vector<unsigned long long> t(64 * 16);
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaMemcpy(t.data(), h->wf_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            const unsigned long long t0 = t[0];
            fprintf(stderr, "[wf trace, us since band 0's loader started; group 0]\n");
            for (int bb = 0; bb < std::min(64, NBP); bb += (bb < 8 ? 1 : 8)) {
                fprintf(stderr, "band %2d:", bb);
                for (int e = 0; e < 13; ++e) fprintf(stderr, " %s=%.1f", names[e], t[bb * 16 + e] ? (double)(t[bb * 16 + e] - t0) / 1e3 : -1.0);
                fprintf(stderr, " | compute warp kcycles: wait_full=%.0f pre=%.0f body=%.0f\n", t[bb * 16 + 13] / 1e3, t[bb * 16 + 14] / 1e3, t[bb * 16 + 15] / 1e3);
            }
        }
        done += kc;
    }
    for (int i = 0; i < nreq; ++i) {
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
    }
    return EQ_OK;
}

// rank r's copy of one of my float arrays (a field or the red-black ping-pong buffer); nullptr outside [0, world)
static int peer_buffer(eq_fluid *h, const float *mine, int r, float **out) {
    *out = nullptr;
    if (r < 0 || r >= h->world) return EQ_OK;
    if (mine == h->rb_tmp) {
        *out = h->peer_tmp[r];
        return EQ_OK;
    }
    const int fi = field_index(h, mine);
    if (fi < 0) return eq_fail(EQ_ERR_INVALID, "peer look-up of an unknown array");
    *out = h->peer_f[r][fi];
    return EQ_OK;
}

// Sliding-window red-black kernel (k_linsolve_rbs.cuh): RS_T iterations per pass, ping-pong between x and rb_tmp.
static int lin_solve_red_black_slide(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    const EqLayout L = h->L;
    const int rows = L.row1 - L.row0;
    // a task (one warp) = a strip of RS_SW columns x a segment of rows; segments as long as possible (each recomputes
    // 2 RS_VH rows) but enough of them to give every SM ~16 warps
    const int nstrips = (L.N + RS_SW - 1) / RS_SW;
    const int want_segs = std::max(1, (16 * h->sm_count + nstrips - 1) / nstrips);
    const int seg_rows = std::min(RS_SEG, std::max(32, (rows + want_segs - 1) / want_segs));
    const int nsegs = (rows + seg_rows - 1) / seg_rows;
    const int grid = (nstrips * nsegs + RS_WARPS - 1) / RS_WARPS;
    for (int i = 0; i < nreq; ++i) {
        const float c_recip = 1.0f / req[i].c;
        float *cur = req[i].x, *other = h->rb_tmp;
        TRY(halo_xchg(h, const_cast<float *>(req[i].x0), RB_H));    // the first / last segment recomputes RS_VH ghost rows
        for (int64_t done = 0; done < iters; done += RS_T) {
            const int it = (int)std::min<int64_t>(RS_T, iters - done);
            TRY(halo_xchg(h, cur, RB_H));                           // ghost rows of the current iterate (also the neighbour barrier)
            EQ_LAUNCH(k_rb_slide, grid, RS_THREADS, 0, h->stream, cur, other, req[i].x0, h->codes, h->chunk_flags, req[i].a,
                      c_recip, req[i].orient, it, L.row0, L.row1, nstrips, nsegs, seg_rows, h->run_if, L);
            TRY(check_launch("k_rb_slide"));
            std::swap(cur, other);
        }
        if (cur != req[i].x) {
            if (h->run_if) {      // the launches above may have been skipped (a == 0 shortcut): then `cur` holds nothing
                const dim3 g((unsigned)((L.P / 4 + 255) / 256), (unsigned)std::min(rows, 1024), 1);
                EQ_LAUNCH(k_copy_rows_if, g, 256, 0, h->stream, req[i].x, cur, h->run_if, L.row0, L.row1, L);
                TRY(check_launch("k_copy_rows_if"));
            } else {
                CU(cudaMemcpyAsync(req[i].x + (size_t)L.row0 * L.P, cur + (size_t)L.row0 * L.P,
                                   (size_t)rows * L.P * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
            }
        }
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
        TRY(halo_xchg(h, req[i].x));
    }
    return EQ_OK;
}

// Streaming red-black kernel (k_rb_stream): RQ_T iterations per pass, rows brought in by bulk copies; what is left of
// `iters` after the last full pass goes through k_rb_slide (2 or 1 iterations per pass).
static int lin_solve_red_black_stream(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    EqLayout L = h->L;
#ifdef EQ_DEBUG_KNOBS
    // timing experiments only (wrong results): pretend the slab has this many rows, to study small slabs on one GPU
    if (debug_knob("EQ_RQ_DEBUG_ROWS")) L.row1 = std::min(L.row1, L.row0 + std::max(64, atoi(getenv("EQ_RQ_DEBUG_ROWS"))));
#endif
    const int rows = L.row1 - L.row0;
    // k_rb_stream: a task (one warp) = a strip of RQ_SW columns x a segment of rows.  Every segment recomputes
    // 2 RQ_VH rows and fills its window (13 rows), so long segments waste less -- but the tasks differ in cost (rows with
    // mirror codes run set_boundaries, ~2x the time) and the hardware balances them only when there are several waves of
    // them: measured at 16384^2 (2368 warps resident), segments of 1171 / 585 / 293 / 195 rows give 3.66 / 3.59 /
    // 3.50 / 3.43 ms per Passive solve and 5.45 / 4.92 / 4.26 / 4.12 ms per AdjustRow solve (16 rectangles); at 4096^2
    // 195 rows is the best as well.  The strips that touch the left / right wall run range tests and
    // set_boundaries on every row (~3x the time per row): quarter-length segments, launched first.
    const int nstrips = (L.N + RQ_SW - 1) / RQ_SW;
    int n_edge = 1;
    for (int sx = nstrips - 1; sx >= 1 && sx * RQ_SW - RQ_HALO + 127 > L.N - 2; --sx) ++n_edge;
    n_edge = std::min(n_edge, nstrips);
    const int ns = nstrips - n_edge;
    const int slots = h->sm_count * RQ_CTAS_PER_SM;
    int nsegs = std::max(1, (rows + RQ_SEG_ROWS / 2) / RQ_SEG_ROWS);
    // about one wave of tasks or less (a row slab, a mid-size grid): shorter tasks, down to 96 rows, as long as they stay
    // within one and a half waves (2048-row slab of 16384 columns: 7 / 11 / 14 / 21 segments -> 0.80 / 0.65 / 0.63 / 0.57 ms)
    if ((long long)std::max(1, ns) * nsegs <= (3LL * slots) / 2)
        nsegs = std::max(nsegs, std::min(rows / 96, (int)((3LL * slots) / 2 / std::max(1, ns))));
    nsegs = env_int("EQ_RQ_SEGS", std::max(1, nsegs));
    const int seg_rows = (rows + nsegs - 1) / nsegs;
    nsegs = (rows + seg_rows - 1) / seg_rows;
    // Wall strips: a range-tested tick with set_boundaries on every row costs ~3.4 fast ticks.  With several waves of
    // tasks they overlap with the rest (a quarter of the segment, at least 64 rows: less recomputation).  When the whole
    // launch is about one wave -- a row slab of a multi-GPU run, a mid-size grid -- the launch lasts as long as its slowest
    // task, so a wall task gets as many rows as make it last as long as an interior task (warm-up rows included), at least 16
    // (measured on a 2048-row slab of 16384 columns: the launch took 0.17 ms whatever the interior segmentation was --
    // the time of a 64 + 37 row wall task).
    int seg_rows_e = std::max(64, seg_rows / 4);
    if ((long long)ns * nsegs <= (3LL * slots) / 2)
        seg_rows_e = std::max(16, (int)((seg_rows + 2 * RQ_VH + 13) * 0.3f) - (2 * RQ_VH + 13));
    seg_rows_e = env_int("EQ_RQ_EDGE_ROWS", std::min(rows, seg_rows_e));
    const int nsegs_e = (rows + seg_rows_e - 1) / seg_rows_e;
    const int grid = n_edge * nsegs_e + ns * nsegs;
    // k_rb_slide for the remainder
    const int nstrips2 = (L.N + RS_SW - 1) / RS_SW;
    const int want_segs = std::max(1, (16 * h->sm_count + nstrips2 - 1) / nstrips2);
    const int seg_rows2 = std::min(RS_SEG, std::max(32, (rows + want_segs - 1) / want_segs));
    const int nsegs2 = (rows + seg_rows2 - 1) / seg_rows2;
    const int grid2 = (nstrips2 * nsegs2 + RS_WARPS - 1) / RS_WARPS;
    for (int i = 0; i < nreq; ++i) {
        const float c_recip = 1.0f / req[i].c;
        float *cur = req[i].x, *other = h->rb_tmp;
        TRY(halo_xchg(h, const_cast<float *>(req[i].x0), RB_H));    // the first / last segment recomputes ghost rows
        int64_t done = 0;
        bool pushed = false;     // the previous k_rb_stream launch wrote my boundary rows into the neighbours' ghost rows
        while (done < iters) {
            // ghost rows of `cur`: copied by the exchange kernel, or already pushed by the pass that produced `cur`
            // (then the exchange is only the neighbour barrier: nrows = 0)
            TRY(halo_xchg(h, cur, pushed ? 0 : RB_H));
            if (iters - done >= RQ_T) {
                if (h->world > 1) {
                    float *pu = nullptr, *pd = nullptr;
                    TRY(peer_buffer(h, other, h->rank - 1, &pu));
                    TRY(peer_buffer(h, other, h->rank + 1, &pd));
                    EQ_LAUNCH(k_rb_stream<true>, grid, RQ_THREADS, RQ_SMEM_BYTES, h->stream, cur, other, req[i].x0, h->codes,
                              h->chunk_flags, req[i].a, c_recip, req[i].orient, L.row0, L.row1, nstrips, n_edge, nsegs, seg_rows,
                              nsegs_e, seg_rows_e, pu, pd, h->run_if, L);
                    pushed = true;
                } else {
                    EQ_LAUNCH(k_rb_stream<false>, grid, RQ_THREADS, RQ_SMEM_BYTES, h->stream, cur, other, req[i].x0, h->codes,
                              h->chunk_flags, req[i].a, c_recip, req[i].orient, L.row0, L.row1, nstrips, n_edge, nsegs, seg_rows,
                              nsegs_e, seg_rows_e, nullptr, nullptr, h->run_if, L);
                }
                TRY(check_launch("k_rb_stream"));
                done += RQ_T;
            } else {
                const int it = (int)std::min<int64_t>(RS_T, iters - done);
                EQ_LAUNCH(k_rb_slide, grid2, RS_THREADS, 0, h->stream, cur, other, req[i].x0, h->codes, h->chunk_flags, req[i].a,
                          c_recip, req[i].orient, it, L.row0, L.row1, nstrips2, nsegs2, seg_rows2, h->run_if, L);
                TRY(check_launch("k_rb_slide"));
                pushed = false;
                done += it;
            }
            std::swap(cur, other);
        }
        if (cur != req[i].x) {
            if (h->run_if) {      // the launches above may have been skipped (a == 0 shortcut): then `cur` holds nothing
                const dim3 g((unsigned)((L.P / 4 + 255) / 256), (unsigned)std::min(rows, 1024), 1);
                EQ_LAUNCH(k_copy_rows_if, g, 256, 0, h->stream, req[i].x, cur, h->run_if, L.row0, L.row1, L);
                TRY(check_launch("k_copy_rows_if"));
            } else {
                CU(cudaMemcpyAsync(req[i].x + (size_t)L.row0 * L.P, cur + (size_t)L.row0 * L.P,
                                   (size_t)rows * L.P * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
            }
        }
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
        TRY(halo_xchg(h, req[i].x));
    }
    return EQ_OK;
}

static int lin_solve_red_black(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    // Grids that fit one SM's shared memory (the reference's 128^2 default scene): one CTA, one launch for all iterations
    {
        const char *rk = getenv("EQ_RB_KERNEL");
        const bool fits = h->world == 1 && (long long)h->L.N * h->L.P <= RBSM_MAX_CELLS && h->L.P / 2 <= RBSM_THREADS &&
                          (h->L.N + RBSM_THREADS / (h->L.P / 2) - 1) / (RBSM_THREADS / (h->L.P / 2)) <= RBSM_MAXR;
        if (fits && (rk ? !strcmp(rk, "small") : true) && iters <= 0x7fffffff) {
            for (int i = 0; i < nreq; ++i) {
                const size_t smem = (size_t)h->L.N * h->L.P * sizeof(float);
                const int rg = RBSM_THREADS / (h->L.P / 2);
                if ((h->L.N + rg - 1) / rg <= 8)
                    EQ_LAUNCH(k_rb_small<8>, 1, RBSM_THREADS, smem, h->stream, req[i].x, req[i].x0, h->codes, req[i].a,
                              1.0f / req[i].c, req[i].orient, (int)iters, h->run_if, h->L);
                else
                    EQ_LAUNCH(k_rb_small<RBSM_MAXR>, 1, RBSM_THREADS, smem, h->stream, req[i].x, req[i].x0, h->codes, req[i].a,
                              1.0f / req[i].c, req[i].orient, (int)iters, h->run_if, h->L);
                TRY(check_launch("k_rb_small"));
            }
            return EQ_OK;
        }
    }
    // Grids of EQ_RB_STREAM_MIN_N columns or more: the streaming kernel (one warp per 104-column strip needs >= ~20
    // strips x a few segments to fill the GPU; below that the register-tile kernel k_rb_reg is faster).
    // EQ_RB_KERNEL=stream|slide|reg|tiled overrides the choice (A/B runs, emulator tests)
    {
        const char *rk = getenv("EQ_RB_KERNEL");
        const bool fits = (unsigned long long)h->L.N * (unsigned long long)h->L.P < (1ull << 32);   // k_rb_stream: 32-bit element offsets
        if (fits && (rk ? !strcmp(rk, "stream") : (h->L.N >= EQ_RB_STREAM_MIN_N))) return lin_solve_red_black_stream(h, req, nreq, iters);
        if (rk && !strcmp(rk, "slide")) return lin_solve_red_black_slide(h, req, nreq, iters);
    }
    const EqLayout L = h->L;
    const int rows = L.row1 - L.row0;
    // EQ_RB_KERNEL=tiled selects the older shared-memory kernel (A/B runs); the default keeps the tile in registers
    static const bool use_tiled = getenv("EQ_RB_KERNEL") && !strcmp(getenv("EQ_RB_KERNEL"), "tiled");
    const dim3 grid_t((unsigned)((L.N + RB_TW - 1) / RB_TW), (unsigned)((rows + RB_TH - 1) / RB_TH), 1);
    // k_rb_reg: tiles on a fixed lattice of RBR_WO x RBR_HO outputs; this rank runs the tile rows that meet its slab
    const int ty0 = L.row0 / RBR_HO, ty1 = (L.row1 - 1) / RBR_HO;
    const dim3 grid_r((unsigned)((L.N + RBR_WO - 1) / RBR_WO), (unsigned)(ty1 - ty0 + 1), 1);
    for (int i = 0; i < nreq; ++i) {
        const float c_recip = 1.0f / req[i].c;
        float *cur = req[i].x, *other = h->rb_tmp;
        // tiles at a slab edge recompute RB_H rows of the neighbour: they need its x0 there
        TRY(halo_xchg(h, const_cast<float *>(req[i].x0), RB_H));
        bool pushed = false;     // the previous k_rb_reg launch wrote my boundary rows into the neighbours' ghost rows
        for (int64_t done = 0; done < iters; done += RB_T) {
            const int it = (int)std::min<int64_t>(RB_T, iters - done);
            // ghost rows of `cur`: copied by the exchange kernel before the first launch, pushed by k_rb_reg itself
            // afterwards (then the exchange is only the neighbour barrier: nrows = 0)
#if RBR_INLINE_BARRIER
            if (!pushed || use_tiled) TRY(halo_xchg(h, cur, RB_H));    // later launches wait on the neighbours' flags themselves
#else
            TRY(halo_xchg(h, cur, pushed ? 0 : RB_H));
#endif
            if (use_tiled) {
                EQ_LAUNCH(k_rb_tiled, grid_t, RB_THREADS, RB_SMEM_BYTES, h->stream, cur, other, req[i].x0, h->codes,
                          h->row_fluid, h->col_fluid, req[i].a, c_recip, req[i].orient, it, L.row0, L.row1, L);
                TRY(check_launch("k_rb_tiled"));
            } else {
                float *pu = nullptr, *pd = nullptr;
                if (h->world > 1) {
                    TRY(peer_buffer(h, other, h->rank - 1, &pu));
                    TRY(peer_buffer(h, other, h->rank + 1, &pd));
                }
#if RBR_INLINE_BARRIER
                RbrSync sy;
                memset(&sy, 0, sizeof(sy));
                if (h->world > 1) {
                    sy.mine = h->sync;
                    sy.up = h->rank > 0 ? h->peer_sync[h->rank - 1] : nullptr;
                    sy.down = h->rank + 1 < h->world ? h->peer_sync[h->rank + 1] : nullptr;
                    sy.epoch = ++h->rb_epoch;
                    sy.wait_epoch = pushed ? sy.epoch - 1 : 0u;
                    sy.error = reinterpret_cast<int *>(h->flags + 1);
                }
                EQ_LAUNCH(k_rb_reg, grid_r, RBR_THREADS, RBR_SMEM_BYTES, h->stream, cur, other, req[i].x0, h->codes,
                          h->chunk_flags, h->row_fluid, h->col_fluid, req[i].a, c_recip, req[i].orient, it, L.row0, L.row1,
                          ty0, pu, pd, h->run_if, L, sy);
#else
                EQ_LAUNCH(k_rb_reg, grid_r, RBR_THREADS, RBR_SMEM_BYTES, h->stream, cur, other, req[i].x0, h->codes,
                          h->chunk_flags, h->row_fluid, h->col_fluid, req[i].a, c_recip, req[i].orient, it, L.row0, L.row1,
                          ty0, pu, pd, h->run_if, L);
#endif
                TRY(check_launch("k_rb_reg"));
                pushed = (h->world > 1);
            }
            std::swap(cur, other);
        }
        if (cur != req[i].x) {
            if (h->run_if) {      // the launches above may have been skipped (a == 0 shortcut): then `cur` holds nothing
                const dim3 g((unsigned)((L.P / 4 + 255) / 256), (unsigned)std::min(rows, 1024), 1);
                EQ_LAUNCH(k_copy_rows_if, g, 256, 0, h->stream, req[i].x, cur, h->run_if, L.row0, L.row1, L);
                TRY(check_launch("k_copy_rows_if"));
            } else {
                CU(cudaMemcpyAsync(req[i].x + (size_t)L.row0 * L.P, cur + (size_t)L.row0 * L.P,
                                   (size_t)rows * L.P * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
            }
        }
        EQ_LAUNCH(k_corners, 1, 32, 0, h->stream, req[i].x, L);
        TRY(check_launch("k_corners"));
        TRY(halo_xchg(h, req[i].x));
    }
    return EQ_OK;
}

static int set_boundaries(eq_fluid *h, int orient, float *x);
static int lin_solve_dispatch(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters);

// a == 0, c == 1 (k_a0_check in k_stencils.cuh): guard, copy, the solver behind the guard's flag, one set_boundaries.
static bool a0_shortcut_applies(const eq_fluid *h, const LinSolveReq &r) {
    static const bool off = getenv("EQ_A0_FASTPATH") && atoi(getenv("EQ_A0_FASTPATH")) == 0;
    return !off && r.a == 0.0f && !std::signbit(r.a) && r.c == 1.0f;
}

static int lin_solve_a0(eq_fluid *h, const LinSolveReq &r, int slot, int64_t iters) {
    const EqLayout L = h->L;
    unsigned *flag = h->a0_flags + slot;
    {
        ProfScope ps(h, CAT_OTHER, 2);
        CU(cudaMemsetAsync(flag, 0, sizeof(unsigned), h->stream));
        const dim3 grid((unsigned)((L.N + 4 * 256 - 1) / (4 * 256)), (unsigned)std::min(L.row1 - L.row0, 2048), 1);
        EQ_LAUNCH(k_a0_check, grid, 256, 0, h->stream, r.x, r.x0, flag, L);
        TRY(check_launch("k_a0_check"));
        if (h->world > 1) {                                   // every rank must come to the same decision
            TRY(need_attached(h));
            EqBarrierArgs a;
            memset(&a, 0, sizeof(a));
            a.sync = h->sync;
            for (int q = 0; q < h->world; ++q) a.peer_sync[q] = h->peer_sync[q];
            a.rank = h->rank;
            a.world = h->world;
            a.epoch = ++h->or_epoch;
            a.error = reinterpret_cast<int *>(h->flags + 1);
            EQ_LAUNCH(k_flag_or_all, 1, 32, 0, h->stream, a, flag);
            TRY(check_launch("k_flag_or_all"));
        }
        EQ_LAUNCH(k_a0_apply, grid, 256, 0, h->stream, r.x, r.x0, flag, L);
        TRY(check_launch("k_a0_apply"));
    }
    h->run_if = flag;
    const int rc = lin_solve_dispatch(h, &r, 1, iters);     // returns at once on the device unless the guard failed
    h->run_if = nullptr;
    TRY(rc);
    TRY(set_boundaries(h, r.orient, r.x));                   // idempotent after a real solve (refreshes the ghost rows it reads)
    return halo_xchg(h, r.x);                                // ghost rows for the stencils that follow
}

static int lin_solve(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    if (iters <= 0) return EQ_OK;   // `for _k in 0..frames` runs zero times
    TRY(ensure_tables(h));
    bool any_a0 = false;
    for (int i = 0; i < nreq; ++i) any_a0 = any_a0 || a0_shortcut_applies(h, req[i]);
    if (any_a0) {
        for (int i = 0; i < nreq; ++i) {
            if (a0_shortcut_applies(h, req[i])) TRY(lin_solve_a0(h, req[i], i & 1, iters));
            else TRY(lin_solve_dispatch(h, req + i, 1, iters));
        }
        return EQ_OK;
    }
    return lin_solve_dispatch(h, req, nreq, iters);
}

static int lin_solve_dispatch(eq_fluid *h, const LinSolveReq *req, int nreq, int64_t iters) {
    const int64_t cells = (int64_t)(h->L.N - 2) * owned_interior_rows(h);
    const int ls_cat = h->run_if ? CAT_OTHER : CAT_LS;      // a solve behind the a == 0 guard normally does no sweeps: not lin_solve time
    if (!h->run_if) h->prof_cell_iters += cells * iters * nreq;
    if (h->prm.mode == EQ_MODE_RED_BLACK) {
        ProfScope ps(h, ls_cat, (int)(((iters + RB_T - 1) / RB_T + 1) * nreq));
        return lin_solve_red_black(h, req, nreq, iters);
    }
    static const bool tb_off = getenv("EQ_LSX_TB") && atoi(getenv("EQ_LSX_TB")) == 0;
    const bool use_tb = (h->world == 1 && TBX_T > 1 && !tb_off);
    const int wave_launches = (int)((iters + LSX_KMAX - 1) / LSX_KMAX) * ((use_tb && !env_int("EQ_TB_BATCH", 0)) ? nreq : 1);
    ProfScope ps(h, ls_cat, wave_launches + nreq);     // + one corner kernel per field
    // EQ_EXACT_KERNEL=tb|wf picks the single-GPU kernel (A/B runs)
    const char *ek = getenv("EQ_EXACT_KERNEL");
    const bool want_wf = ek ? !strcmp(ek, "wf") : (EQ_DEFAULT_EXACT_WF != 0);
    if (h->world == 1 && want_wf && !tb_off) return lin_solve_exact_wf(h, req, nreq, iters);
    if (use_tb) return lin_solve_exact_tb(h, req, nreq, iters);
    // row slabs: the slab kernel, or every rank solving the whole grid (faster on 2 GPUs; EQ_EXACT_PLAN=slab|replica)
    const char *plan = getenv("EQ_EXACT_PLAN");
    const bool replica = plan ? !strcmp(plan, "replica") : (h->world > 1 && h->world <= EQ_EXACT_REPLICA_MAX_WORLD);
    if (h->world > 1 && replica && TBX_T > 1 && !tb_off) return lin_solve_exact_replica(h, req, nreq, iters);
    return lin_solve_exact(h, req, nreq, iters);
}

// diffuse (fluid.rs:276-298): a = dt * diff * (N-2) * (N-2), c = 1 + 4a
static LinSolveReq diffuse_req(const eq_fluid *h, int orient, float *x, const float *x0, float diff) {
    const float sf = (float)(h->L.N - 2);
    float a = h->prm.delta_t * diff;
    a = a * sf;
    a = a * sf;
    return LinSolveReq{orient, x, x0, a, 1.0f + 4.0f * a};
}

// project (fluid.rs:330-375)
static int project(eq_fluid *h, float *vx, float *vy, float *p, float *div, int64_t iters) {
    const EqLayout L = h->L;
    TRY(halo_xchg(h, vy));                                             // the stencil reads vy[j-1], vy[j+1]
    {
        ProfScope ps(h, CAT_PROJ, 1);
        EQ_LAUNCH(k_divergence, stencil_grid(h, owned_interior_rows(h)), EQ_ST_THREADS, 0, h->stream, vx, vy, div, p, L);
        TRY(check_launch("k_divergence"));
    }
    TRY(set_boundaries(h, EQ_PASSIVE, div));                           // :351
    TRY(set_boundaries(h, EQ_PASSIVE, p));                             // :352
    LinSolveReq r{EQ_PASSIVE, p, div, 1.0f, 4.0f};
    TRY(lin_solve(h, &r, 1, iters));                                   // :353-362
    {
        ProfScope ps(h, CAT_PROJ, 1);
        EQ_LAUNCH(k_gradient, stencil_grid(h, owned_interior_rows(h)), EQ_ST_THREADS, 0, h->stream, vx, vy, p, L);
        TRY(check_launch("k_gradient"));
    }
    TRY(set_boundaries(h, EQ_ADJUST_ROW, vx));                         // :373
    TRY(set_boundaries(h, EQ_ADJUST_COLUMN, vy));                      // :374
    return EQ_OK;
}

// advect (fluid.rs:378-432) for one field, or for the velocity pair that shares its back-trace
static int advect(eq_fluid *h, int orientA, float *dA, const float *d0A, int orientB, float *dB,
                  const float *d0B, const float *vx, const float *vy) {
    const EqLayout L = h->L;
    TRY(barrier_all(h));     // the back-trace may sample any rank's rows of d0: everybody must have produced them
    {
        ProfScope ps(h, CAT_ADV, 1);
        const EqPeerTable tA = peer_table(h, d0A), tB = peer_table(h, d0B);
        const int rows = owned_interior_rows(h);
        const float dt = h->prm.delta_t;
        if (dB && h->world > 1) EQ_LAUNCH((k_advect<2, true>), rows, EQ_ADV_THREADS, 16, h->stream, dA, tA, dB, tB, vx, vy, dt, L);
        else if (dB) EQ_LAUNCH((k_advect<2, false>), rows, EQ_ADV_THREADS, 16, h->stream, dA, tA, dB, tB, vx, vy, dt, L);
        else if (h->world > 1) EQ_LAUNCH((k_advect<1, true>), rows, EQ_ADV_THREADS, 16, h->stream, dA, tA, nullptr, tB, vx, vy, dt, L);
        else EQ_LAUNCH((k_advect<1, false>), rows, EQ_ADV_THREADS, 16, h->stream, dA, tA, nullptr, tB, vx, vy, dt, L);
        TRY(check_launch("k_advect"));
    }
    TRY(barrier_all(h));     // ... and nobody may overwrite d0 while a neighbour still samples it
    TRY(set_boundaries(h, orientA, dA));                               // :431
    if (dB) TRY(set_boundaries(h, orientB, dB));
    return EQ_OK;
}

static int64_t gs_iters(const eq_fluid *h) {
    return h->prm.gs_iterations ? h->prm.gs_iterations : h->prm.frames;   // fluid.rs:445 (quirk Q1)
}

// Fluid::step (fluid.rs:437-524)
static int step_once(eq_fluid *h) {
    float *density = h->f[EQ_F_DENSITY], *vx = h->f[EQ_F_VX], *vy = h->f[EQ_F_VY];
    float *vx0 = h->f[EQ_F_VX0], *vy0 = h->f[EQ_F_VY0], *scratch = h->f[EQ_F_SCRATCH];
    const int64_t k = gs_iters(h);
    // :438-457  the two velocity diffusions are independent: one wavefront launch
    LinSolveReq d2[2] = {diffuse_req(h, EQ_ADJUST_ROW, vx0, vx, h->prm.viscosity),
                         diffuse_req(h, EQ_ADJUST_COLUMN, vy0, vy, h->prm.viscosity)};
    TRY(lin_solve(h, d2, 2, k));
    TRY(project(h, vx0, vy0, vx, vy, k));                              // :459-467
    TRY(advect(h, EQ_ADJUST_ROW, vx, vx0, EQ_ADJUST_COLUMN, vy, vy0, vx0, vy0));   // :469-489
    TRY(project(h, vx, vy, vx0, vy0, k));                              // :491-499
    LinSolveReq dd = diffuse_req(h, EQ_PASSIVE, scratch, density, h->prm.diffusion);
    TRY(lin_solve(h, &dd, 1, k));                                      // :501-510
    TRY(advect(h, EQ_PASSIVE, density, scratch, 0, nullptr, nullptr, vx, vy));     // :512-521
    {
        ProfScope ps(h, CAT_OTHER, 1);
        const int r0 = std::max(0, h->L.row0 - 1), r1 = std::min(h->L.N, h->L.row1 + 1);   // slab + ghost rows
        CU(cudaMemcpyAsync(scratch + (size_t)r0 * h->L.P, density + (size_t)r0 * h->L.P,
                           (size_t)(r1 - r0) * h->L.P * sizeof(float), cudaMemcpyDeviceToDevice,
                           h->stream));                                // :523
    }
    h->prof_steps += 1;
    return EQ_OK;
}

// ---------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------
static int validate_params(const EqParams *p) {
    if (!p) return eq_fail(EQ_ERR_INVALID, "null params");
    if (p->size < 20 || p->size > 32768)
        return eq_fail(EQ_ERR_INVALID, "size %u out of range [20, 32768] (init_density underflows below 20, fluid.rs:534)",
                       p->size);
    if (p->frames < 0 || p->gs_iterations < 0) return eq_fail(EQ_ERR_INVALID, "negative frames / gs_iterations");
    if (p->mode != EQ_MODE_EXACT && p->mode != EQ_MODE_RED_BLACK) return eq_fail(EQ_ERR_INVALID, "unknown mode %d", p->mode);
    if (p->world > EQ_MAX_RANKS) return eq_fail(EQ_ERR_INVALID, "world %d > %d", p->world, EQ_MAX_RANKS);
    if (p->world > 1 && (p->rank < 0 || p->rank >= p->world)) return eq_fail(EQ_ERR_INVALID, "rank %d not in [0,%d)", p->rank, p->world);
    if (p->world > 1 && ((int)p->size - 2 + 31) / 32 < p->world)
        return eq_fail(EQ_ERR_INVALID, "grid too small for %d row slabs (one 32-row band per rank at least)", p->world);
    return EQ_OK;
}

static int alloc_handle(const EqParams *params, eq_fluid **out) {
    TRY(validate_params(params));
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return eq_fail(EQ_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (params->device < 0 || params->device >= ndev) return eq_fail(EQ_ERR_INVALID, "device %d not in [0,%d)", params->device, ndev);
    CU(cudaSetDevice(params->device));
    eq_fluid *h = new (std::nothrow) eq_fluid();
    if (!h) return eq_fail(EQ_ERR_NOMEM, "host allocation failed");
    memset(h->f, 0, sizeof(h->f));
    h->prm = *params;
    h->dev = params->device;
    h->L.N = (int)params->size;
    h->L.P = ((int)params->size + 31) / 32 * 32;
    h->L.rows = (int)params->size + EQ_ROW_PAD;
    {
        const int NBt = (h->L.N - 2 + 31) / 32;
        h->world = std::max(1, (int)params->world);
        h->rank = h->world > 1 ? params->rank : 0;
        h->b_lo = (int)((int64_t)h->rank * NBt / h->world);
        h->b_hi = (int)((int64_t)(h->rank + 1) * NBt / h->world);
        h->L.row0 = h->rank == 0 ? 0 : 1 + 32 * h->b_lo;
        h->L.row1 = h->rank == h->world - 1 ? h->L.N : 1 + 32 * h->b_hi;
        h->L.rank = h->rank;
        h->L.world = h->world;
    }
    h->spans = new std::vector<ProfSpan>();
    h->ev_pool = new std::vector<cudaEvent_t>();
    h->job_tables = new std::map<int, uint32_t *>();
    h->mask_dirty = true;
    *out = h;   // so that the caller can destroy a half-built handle
    CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->dev));
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    const size_t elems = field_elems(h);
    for (int i = 0; i < 6; ++i) {
        CU(cudaMalloc(&h->f[i], elems * sizeof(float)));
        CU(cudaMemsetAsync(h->f[i], 0, elems * sizeof(float), h->stream));
    }
    CU(cudaMalloc(&h->rb_tmp, elems * sizeof(float)));
    CU(cudaMemsetAsync(h->rb_tmp, 0, elems * sizeof(float), h->stream));
    CU(cudaMalloc(&h->cells, elems));
    CU(cudaMemsetAsync(h->cells, 0, elems, h->stream));
    CU(cudaMalloc(&h->codes, elems));
    CU(cudaMemsetAsync(h->codes, 0, elems, h->stream));
    CU(cudaMalloc(&h->row_fluid, h->L.N));
    CU(cudaMalloc(&h->col_fluid, h->L.P));
    CU(cudaMalloc(&h->counts, 4 * sizeof(unsigned)));
    CU(cudaMalloc(&h->a0_flags, 2 * sizeof(unsigned)));
    CU(cudaMalloc(&h->chunk_flags, 2 * (size_t)((h->L.N - 2 + 31) / 32) * ((h->L.N + EQ_LSX_CW - 1) / EQ_LSX_CW)));
    {
        const int NBP = (h->L.N - 2 + TBX_SK + 31) / 32;
        CU(cudaMalloc(&h->chunk_flags_tb, 2 * (size_t)NBP * ((h->L.N + EQ_LSX_CW - 1) / EQ_LSX_CW)));
        for (int i = 0; i < 2; ++i) {
            const size_t nraw = (size_t)TBX_T * NBP * h->L.P, nedge = (size_t)std::max(1, TBX_T - 1) * NBP * 2 * h->L.P;
            CU(cudaMalloc(&h->tb_raw[i], nraw * sizeof(float)));
            CU(cudaMemsetAsync(h->tb_raw[i], 0, nraw * sizeof(float), h->stream));
            CU(cudaMalloc(&h->tb_edge[i], nedge * sizeof(float)));
            CU(cudaMemsetAsync(h->tb_edge[i], 0, nedge * sizeof(float), h->stream));
        }
    }
    {
        const int NBPw = (h->L.N + WF_SK + 31) / 32, NCw = h->L.P / WF_CW;
        CU(cudaMalloc(&h->wf_flags, 3 * (size_t)NBPw * NCw));
        for (int i = 0; i < 2; ++i) {
            const size_t nraw = (size_t)WF_T * NBPw * h->L.P, nedge = (size_t)(WF_T - 1) * NBPw * 2 * h->L.P;
            CU(cudaMalloc(&h->wf_raw[i], nraw * sizeof(float)));
            CU(cudaMemsetAsync(h->wf_raw[i], 0, nraw * sizeof(float), h->stream));
            CU(cudaMalloc(&h->wf_edge[i], nedge * sizeof(float)));
            CU(cudaMemsetAsync(h->wf_edge[i], 0, nedge * sizeof(float), h->stream));
        }
    }
    const int NB = (h->L.N - 2 + 31) / 32;
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&h->raw[i], (size_t)NB * h->L.P * sizeof(float)));
        CU(cudaMemsetAsync(h->raw[i], 0, (size_t)NB * h->L.P * sizeof(float), h->stream));
    }
    h->flags_words = 8 + 2 * (size_t)LSX_KMAX * NB;
    CU(cudaMalloc(&h->flags, h->flags_words * sizeof(unsigned)));
    CU(cudaMemsetAsync(h->flags, 0, h->flags_words * sizeof(unsigned), h->stream));
    CU(cudaMalloc(&h->sync, EQ_SYNC_WORDS * sizeof(unsigned)));
    CU(cudaMemsetAsync(h->sync, 0, EQ_SYNC_WORDS * sizeof(unsigned), h->stream));
    for (int i = 0; i < 6; ++i) h->peer_f[h->rank][i] = h->f[i];
    h->peer_raw[h->rank][0] = h->raw[0];
    h->peer_raw[h->rank][1] = h->raw[1];
    h->peer_flags[h->rank] = h->flags;
    h->peer_tmp[h->rank] = h->rb_tmp;
    h->peer_sync[h->rank] = h->sync;
    CU(cudaFuncSetAttribute(k_linsolve_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LSX_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rb_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rb_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RBR_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rb_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RQ_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rb_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RQ_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rb_small<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RBSM_MAX_CELLS * sizeof(float))));
    CU(cudaFuncSetAttribute(k_rb_small<RBSM_MAXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RBSM_MAX_CELLS * sizeof(float))));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linsolve_exact, LSX_THREADS, LSX_SMEM_BYTES));
    h->lsx_ctas = std::max(1, per_sm) * h->sm_count;
    {
        CU(cudaFuncSetAttribute(k_linsolve_tb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TBX_SMEM_BYTES));
        int tb_per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tb_per_sm, k_linsolve_tb, TBX_THREADS, TBX_SMEM_BYTES));
        h->tb_ctas = std::max(1, tb_per_sm) * h->sm_count;
    }
    {
        CU(cudaFuncSetAttribute(k_linsolve_wf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES));
        int wf_per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wf_per_sm, k_linsolve_wf, WF_THREADS, WF_SMEM_BYTES));
        h->wf_ctas = std::max(1, wf_per_sm) * h->sm_count;
        if (const char *e = getenv("EQ_WF_CTAS_PER_SM")) h->wf_ctas = std::min(h->wf_ctas, std::max(1, atoi(e)) * h->sm_count);
    }
    if (const char *e = getenv("EQ_LSX_CTAS_PER_SM")) h->lsx_ctas = std::max(1, std::min(per_sm, atoi(e))) * h->sm_count;
    if (const char *e = getenv("EQ_LSX_CTAS_PER_SM")) h->tb_ctas = std::min(h->tb_ctas, std::max(1, atoi(e)) * h->sm_count);
    if (getenv("EQ_LSX_TRACE")) {
        CU(cudaMalloc(&h->lsx_trace, 4 * 8 * 128 * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(h->lsx_trace, 0, 4 * 8 * 128 * sizeof(unsigned long long), h->stream));
    }
    if (getenv("EQ_LSX_STATS")) {
        CU(cudaMalloc(&h->lsx_stats, 16 * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(h->lsx_stats, 0, 16 * sizeof(unsigned long long), h->stream));
    }
    return EQ_OK;
}

static int clamp_i64(int64_t v, int64_t lo, int64_t hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); }

static int launch_add_rect(eq_fluid *h, float *f, int x0, int y0, int x1, int y1, float amount) {
    if (x1 <= x0 || y1 <= y0) return EQ_OK;
    dim3 grid((unsigned)((x1 - x0 + 255) / 256), (unsigned)(y1 - y0), 1);
    EQ_LAUNCH(k_add_rect, grid, 256, 0, h->stream, f, x0, y0, x1, y1, amount, h->L);
    return check_launch("k_add_rect");
}

static int launch_cells_rect(eq_fluid *h, int x0, int y0, int x1, int y1, uint8_t v) {
    if (x1 <= x0 || y1 <= y0) return EQ_OK;
    dim3 grid((unsigned)((x1 - x0 + 255) / 256), (unsigned)(y1 - y0), 1);
    EQ_LAUNCH(k_set_cells_rect, grid, 256, 0, h->stream, h->cells, x0, y0, x1, y1, v, h->L);
    h->mask_dirty = true;
    return check_launch("k_set_cells_rect");
}

static int init_walls(eq_fluid *h) {   // fluid.rs:552-570
    const int N = h->L.N;
    TRY(launch_cells_rect(h, 0, 0, N, 1, 1));
    TRY(launch_cells_rect(h, 0, N - 1, N, N, 1));
    TRY(launch_cells_rect(h, 0, 0, 1, N, 1));
    TRY(launch_cells_rect(h, N - 1, 0, N, N, 1));
    return EQ_OK;
}

static int init_default(eq_fluid *h) {   // fluid.rs:602-606
    const int N = h->L.N, c = N / 2;
    ProfScope ps(h, CAT_OTHER, 8);
    TRY(launch_add_rect(h, h->f[EQ_F_VX], 0, 0, N, N, 1.0f));                       // :542-548
    TRY(launch_add_rect(h, h->f[EQ_F_VY], 0, 0, N, N, 1.0f));
    TRY(launch_add_rect(h, h->f[EQ_F_DENSITY], c - 10, c - 10, c + 11, c + 11, 0.9f));   // :534-538
    TRY(launch_add_rect(h, h->f[EQ_F_SCRATCH], c - 10, c - 10, c + 11, c + 11, 0.9f));
    return init_walls(h);
}

// ---------------------------------------------------------------------------
// exported API
// ---------------------------------------------------------------------------
extern "C" {

const char *eq_last_error(void) { return g_err; }
int eq_abi_version(void) { return EQ_ABI_VERSION; }

// A freshly leased box can answer the very first CUDA call of its life with a transient error
// (cudaErrorSystemNotReady while the fabric manager is still bringing NVSwitch up,
// cudaErrorDevicesUnavailable, ...).  Retry for EQ_DEVICE_WAIT_S seconds (default 30) before
// giving up, and keep the reason in eq_last_error() either way so that callers can print it.
int eq_device_count(void) {
    int wait_s = env_int("EQ_DEVICE_WAIT_S", 30);
    int n = 0;
    cudaError_t e = cudaSuccess;
    for (int tries = 0;; tries++) {
        n = 0;
        e = cudaGetDeviceCount(&n);
        if (e == cudaSuccess && n > 0) {
            g_err[0] = 0;
            return n;
        }
        if (e != cudaSuccess) cudaGetLastError();
        // "no device" is final on a box without a GPU: do not make CPU-only callers wait
        if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) break;
        if (tries * 250 >= wait_s * 1000) break;
        usleep(250 * 1000);
    }
    if (e != cudaSuccess)
        eq_fail(EQ_ERR_CUDA, "cudaGetDeviceCount failed: %s (%s)", cudaGetErrorString(e), cudaGetErrorName(e));
    else
        eq_fail(EQ_ERR_CUDA, "cudaGetDeviceCount succeeded but reports 0 devices (CUDA_VISIBLE_DEVICES=%s)",
                getenv("CUDA_VISIBLE_DEVICES") ? getenv("CUDA_VISIBLE_DEVICES") : "<unset>");
    return 0;
}

int eq_destroy(eq_fluid *h) {
    if (!h) return EQ_OK;
    cudaSetDevice(h->dev);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    for (int i = 0; i < 6; ++i) cudaFree(h->f[i]);
    cudaFree(h->cells);
    cudaFree(h->rb_tmp);
    cudaFree(h->codes);
    cudaFree(h->row_fluid);
    cudaFree(h->col_fluid);
    cudaFree(h->counts);
    cudaFree(h->a0_flags);
    cudaFree(h->chunk_flags);
    cudaFree(h->chunk_flags_tb);
    cudaFree(h->wf_flags);
    cudaFree(h->wf_dbg);
    cudaFree(h->wf_trace);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->tb_raw[i]);
        cudaFree(h->tb_edge[i]);
        cudaFree(h->wf_raw[i]);
        cudaFree(h->wf_edge[i]);
    }
    cudaFree(h->lsx_stats);
    cudaFree(h->lsx_trace);
    cudaFree(h->lsx_jobtimes);
    cudaFree(h->row_list);
    cudaFree(h->col_list);
    cudaFree(h->raw[0]);
    cudaFree(h->raw[1]);
    for (int r = 0; r < EQ_MAX_RANKS; ++r)
        if (h->peer_ipc[r]) {
            for (int i = 0; i < 6; ++i) cudaIpcCloseMemHandle(h->peer_f[r][i]);
            cudaIpcCloseMemHandle(h->peer_raw[r][0]);
            cudaIpcCloseMemHandle(h->peer_raw[r][1]);
            cudaIpcCloseMemHandle(h->peer_flags[r]);
            cudaIpcCloseMemHandle(h->peer_sync[r]);
            cudaIpcCloseMemHandle(h->peer_tmp[r]);
        }
    cudaFree(h->flags);
    cudaFree(h->sync);
    cudaFree(h->l2buf);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    for (int i = 0; i < EQ_SNAPSHOT_SLOTS; ++i) {
        cudaFree(h->snap_buf[i]);
        if (h->snap_ready[i]) cudaEventDestroy(h->snap_ready[i]);
        if (h->snap_done[i]) cudaEventDestroy(h->snap_done[i]);
    }
    if (h->job_tables) {
        for (auto &kv : *h->job_tables) cudaFree(kv.second);
        delete h->job_tables;
    }
    if (h->spans) {
        for (auto &s : *h->spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        delete h->spans;
    }
    if (h->ev_pool) {
        for (auto e : *h->ev_pool) cudaEventDestroy(e);
        delete h->ev_pool;
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return EQ_OK;
}

int eq_create(const EqParams *params, eq_fluid **out) {
    if (!out) return eq_fail(EQ_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int r = alloc_handle(params, out);
    if (r == EQ_OK) r = init_default(*out);                            // Fluid::new calls init(), fluid.rs:108
    if (r != EQ_OK) {
        eq_destroy(*out);
        *out = nullptr;
    }
    return r;
}

int eq_clone(eq_fluid *h, eq_fluid **out) {
    NEED(h);
    if (!out) return eq_fail(EQ_ERR_INVALID, "null out pointer");
    if (h->world > 1) return eq_fail(EQ_ERR_STATE, "eq_clone of a row-slab handle: download the owned rows instead");
    *out = nullptr;
    int r = alloc_handle(&h->prm, out);
    if (r != EQ_OK) {
        eq_destroy(*out);
        *out = nullptr;
        return r;
    }
    eq_fluid *c = *out;
    const size_t elems = field_elems(h);
    CU(cudaStreamSynchronize(c->stream));   // the clone's memsets ran on its own stream
    for (int i = 0; i < 6; ++i)
        CU(cudaMemcpyAsync(c->f[i], h->f[i], elems * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(c->cells, h->cells, elems, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    c->mask_dirty = true;
    return EQ_OK;
}

int eq_init_default(eq_fluid *h) {
    NEED(h);
    return init_default(h);
}

int eq_add_density(eq_fluid *h, uint32_t x, uint32_t y, float amount) {
    NEED(h);
    const int N = h->L.N;
    const unsigned o = (unsigned)std::min<uint32_t>(x, N - 1) + (unsigned)std::min<uint32_t>(y, N - 1) * (unsigned)h->L.P;
    ProfScope ps(h, CAT_OTHER, 1);
    EQ_LAUNCH(k_add_source, 1, 32, 0, h->stream, h->f[EQ_F_DENSITY], h->f[EQ_F_SCRATCH], h->f[EQ_F_VX], h->f[EQ_F_VY], o,
                                          amount, 0.f, 0.f, 1);
    return check_launch("k_add_source");
}

int eq_add_velocity(eq_fluid *h, uint32_t x, uint32_t y, float ax, float ay) {
    NEED(h);
    const int N = h->L.N;
    const unsigned o = (unsigned)std::min<uint32_t>(x, N - 1) + (unsigned)std::min<uint32_t>(y, N - 1) * (unsigned)h->L.P;
    ProfScope ps(h, CAT_OTHER, 1);
    EQ_LAUNCH(k_add_source, 1, 32, 0, h->stream, h->f[EQ_F_DENSITY], h->f[EQ_F_SCRATCH], h->f[EQ_F_VX], h->f[EQ_F_VY], o,
                                          0.f, ax, ay, 2);
    return check_launch("k_add_source");
}

int eq_rect_valid(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t size) {
    return x0 != x1 && y0 != y1 && x0 < x1 && y0 < y1 && x0 < size && y0 < size && x1 < size && y1 < size;
}

int eq_fill_rect(eq_fluid *h, int64_t x0, int64_t y0, int64_t x1, int64_t y1) {
    NEED(h);
    if (x0 >= x1 || y0 >= y1) return EQ_OK;                            // empty Rust ranges
    const int N = h->L.N;
    // {clamp(x) : x in [x0,x1)} is the contiguous range [clamp(x0), clamp(x1-1)]
    const int cx0 = clamp_i64(x0, 0, N - 1), cx1 = clamp_i64(x1 - 1, 0, N - 1) + 1;
    const int cy0 = clamp_i64(y0, 0, N - 1), cy1 = clamp_i64(y1 - 1, 0, N - 1) + 1;
    ProfScope ps(h, CAT_OTHER, 1);
    return launch_cells_rect(h, cx0, cy0, cx1, cy1, 1);
}

int eq_reset_walls(eq_fluid *h) {
    NEED(h);
    ProfScope ps(h, CAT_OTHER, 5);
    TRY(launch_cells_rect(h, 0, 0, h->L.N, h->L.N, 0));
    return init_walls(h);
}

int eq_set_params(eq_fluid *h, const EqParams *p) {
    NEED(h);
    if (!p) return eq_fail(EQ_ERR_INVALID, "null params");
    if (p->size != h->prm.size) return eq_fail(EQ_ERR_INVALID, "size is fixed at creation (the reference builds a new Fluid, renderer.rs:145-149)");
    if (p->frames < 0 || p->gs_iterations < 0) return eq_fail(EQ_ERR_INVALID, "negative frames / gs_iterations");
    if (p->mode != EQ_MODE_EXACT && p->mode != EQ_MODE_RED_BLACK) return eq_fail(EQ_ERR_INVALID, "unknown mode %d", p->mode);
    h->prm.delta_t = p->delta_t;
    h->prm.frames = p->frames;
    h->prm.gs_iterations = p->gs_iterations;
    h->prm.diffusion = p->diffusion;
    h->prm.viscosity = p->viscosity;
    h->prm.mode = p->mode;
    return EQ_OK;
}

int eq_get_params(eq_fluid *h, EqParams *out) {
    if (!h || !out) return eq_fail(EQ_ERR_INVALID, "null argument");
    *out = h->prm;
    return EQ_OK;
}

int eq_step(eq_fluid *h) {
    NEED(h);
    return step_once(h);
}

int eq_step_n(eq_fluid *h, int64_t n, const EqSource *sources, int64_t n_sources) {
    NEED(h);
    if (n < 0 || n_sources < 0 || (n_sources > 0 && !sources)) return eq_fail(EQ_ERR_INVALID, "bad step_n arguments");
    int64_t s = 0;
    for (int64_t fr = 0; fr < n; ++fr) {
        while (s < n_sources && sources[s].frame <= fr) {
            if (sources[s].frame == fr) {
                const int N = h->L.N;
                const unsigned o = (unsigned)std::min<uint32_t>(sources[s].x, N - 1) +
                                   (unsigned)std::min<uint32_t>(sources[s].y, N - 1) * (unsigned)h->L.P;
                const int what = (sources[s].d_density != 0.f ? 1 : 0) | 2;
                ProfScope ps(h, CAT_OTHER, 1);
                EQ_LAUNCH(k_add_source, 1, 32, 0, h->stream, h->f[EQ_F_DENSITY], h->f[EQ_F_SCRATCH], h->f[EQ_F_VX],
                                                      h->f[EQ_F_VY], o, sources[s].d_density, sources[s].d_vx,
                                                      sources[s].d_vy, what);
                TRY(check_launch("k_add_source"));
            }
            ++s;
        }
        TRY(step_once(h));
    }
    return EQ_OK;
}

// add_noise on the device: the impulse of Philox counter noise->first_frame
int eq_add_noise(eq_fluid *h, const EqNoise *noise) {
    NEED(h);
    if (!noise) return eq_fail(EQ_ERR_INVALID, "null noise parameters");
    const uint64_t frame = noise->first_frame;
    ProfScope ps(h, CAT_OTHER, 1);
    EQ_LAUNCH(k_add_noise, 1, 32, 0, h->stream, h->f[EQ_F_VX], h->f[EQ_F_VY], (uint32_t)noise->seed,
              (uint32_t)(noise->seed >> 32), (uint32_t)frame, (uint32_t)(frame >> 32), noise->cos_t, noise->sin_t,
              noise->gain, h->L);
    return check_launch("k_add_noise");
}

// n x { add_noise; step }: frame f of this call gets the impulse of Philox counter first_frame + f
int eq_step_n_noise(eq_fluid *h, int64_t n, const EqNoise *noise) {
    NEED(h);
    if (n < 0 || !noise) return eq_fail(EQ_ERR_INVALID, "bad step_n_noise arguments");
    EqNoise nz = *noise;
    for (int64_t fr = 0; fr < n; ++fr, ++nz.first_frame) {
        TRY(eq_add_noise(h, &nz));
        TRY(step_once(h));
    }
    return EQ_OK;
}

static void dump_lsx_stats(eq_fluid *h) {
    if (h->lsx_jobtimes && h->lsx_jobtimes_n) {
        std::vector<unsigned long long> v(h->lsx_jobtimes_n);
        if (cudaMemcpy(v.data(), h->lsx_jobtimes, v.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (FILE *f = fopen(getenv("EQ_LSX_JOBTIMES"), "wb")) {
                fwrite(v.data(), sizeof(unsigned long long), v.size(), f);
                fclose(f);
            }
        }
    }
    if (h->lsx_trace) {
        static unsigned long long tr[4 * 8 * 128];
        if (cudaMemcpy(tr, h->lsx_trace, sizeof(tr), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull;
            for (auto v : tr) if (v && v < t0) t0 = v;
            const char *ev[8] = {"slot_free", "load_issued", "macro_begin", "macro_end", "done_seen", "stored", "release_begin", "released"};
            for (int b = 0; b < 4; ++b)
                for (int e = 0; e < 8; ++e) {
                    fprintf(stderr, "[trace b=%d %-13s]", b, ev[e]);
                    for (int q = 0; q < 40; ++q) {
                        const unsigned long long v = tr[(b * 8 + e) * 128 + q];
                        if (v) fprintf(stderr, " %6.1f", (v - t0) / 1e3); else fprintf(stderr, "      -");
                    }
                    fprintf(stderr, "\n");
                }
            cudaMemset(h->lsx_trace, 0, sizeof(tr));
        }
    }
    if (!h->lsx_stats) return;
    unsigned long long v[16];
    if (cudaMemcpy(v, h->lsx_stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    cudaMemset(h->lsx_stats, 0, sizeof(v));
    const char *names[16] = {"compute.wait_full", "compute.fast", "compute.coded", "compute.edge", "n.fast", "n.coded",
                             "n.edge", "loader.wait_free", "loader.wait_flags", "n.releases", "storer.wait_done", "storer.store",
                             "publisher.release", "-", "-", "-"};
    fprintf(stderr, "[lsx stats, Mcycles summed over jobs]");
    for (int i = 0; i < 13; ++i)
        if (names[i][0] != '-') fprintf(stderr, " %s=%.1f", names[i], ((i >= 4 && i <= 6) || i == 9) ? (double)v[i] : v[i] / 1e6);
    fprintf(stderr, "\n");
}

static int check_device_error(eq_fluid *h) {
    int err = 0;
    CU(cudaMemcpyAsync(&err, h->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (err != 0 && h->wf_dbg) {
        std::vector<unsigned> d((size_t)h->wf_dbg_ctas * 32);
        cudaMemcpy(d.data(), h->wf_dbg, d.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[wf debug] cta: role(job g.b, index, state) for compute / loader / storer / publisher; state 9 = done\n");
        for (int c = 0; c < h->wf_dbg_ctas; ++c) {
            const unsigned *w = d.data() + (size_t)c * 32;
            fprintf(stderr, "cta %4d:", c);
            for (int r = 0; r < 4; ++r) fprintf(stderr, "  %u.%u i=%u s=%u |", w[8 * r] >> 16, w[8 * r] & 0xffffu, w[8 * r + 1], w[8 * r + 2]);
            fprintf(stderr, "\n");
        }
    }
    if (err != 0) return eq_fail(EQ_ERR_TIMEOUT, "wavefront watchdog fired (code %d): a lin_solve job waited too long", err);
    return EQ_OK;
}

int eq_sync(eq_fluid *h) {
    NEED(h);
    CU(cudaStreamSynchronize(h->stream));
    dump_lsx_stats(h);
    return check_device_error(h);
}

static int field_ptr(eq_fluid *h, int field, void **ptr, size_t *esize) {
    if (field >= 0 && field < 6) {
        *ptr = h->f[field];
        *esize = sizeof(float);
        return EQ_OK;
    }
    if (field == EQ_F_CELLS) {
        *ptr = h->cells;
        *esize = 1;
        return EQ_OK;
    }
    return eq_fail(EQ_ERR_INVALID, "unknown field id %d", field);
}

int eq_upload_rows(eq_fluid *h, int field, uint32_t row_begin, uint32_t n_rows, const void *host) {
    NEED(h);
    void *d;
    size_t es;
    TRY(field_ptr(h, field, &d, &es));
    const uint32_t N = (uint32_t)h->L.N;
    if (!host || row_begin > N || n_rows > N - row_begin) return eq_fail(EQ_ERR_INVALID, "bad row range");
    if (n_rows == 0) return EQ_OK;
    if (field == EQ_F_CELLS) {
        const uint8_t *c = static_cast<const uint8_t *>(host);
        for (uint32_t r = 0; r < n_rows; ++r) {
            const uint32_t j = row_begin + r;
            const uint8_t *row = c + (size_t)r * N;
            for (uint32_t i = 0; i < N; ++i)
                if (row[i] > 1) return eq_fail(EQ_ERR_INVALID, "cells_type must be 0 (NoWall) or 1 (DefaultWall)");
            const bool frame_row = (j == 0 || j == N - 1);
            if (frame_row) {
                for (uint32_t i = 0; i < N; ++i)
                    if (!row[i]) return eq_fail(EQ_ERR_INVALID, "frame cell (%u,%u) must be DefaultWall (init_walls, fluid.rs:552-570)", i, j);
            } else if (!row[0] || !row[N - 1]) {
                return eq_fail(EQ_ERR_INVALID, "frame cells of row %u must be DefaultWall (init_walls, fluid.rs:552-570)", j);
            }
        }
        h->mask_dirty = true;
    }
    ProfScope ps(h, CAT_OTHER, 0);
    CU(cudaMemcpy2DAsync(static_cast<char *>(d) + (size_t)row_begin * h->L.P * es, (size_t)h->L.P * es, host,
                         (size_t)N * es, (size_t)N * es, n_rows, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // the host buffer is only borrowed for the call
    return EQ_OK;
}

int eq_download_rows(eq_fluid *h, int field, uint32_t row_begin, uint32_t n_rows, void *host) {
    NEED(h);
    void *d;
    size_t es;
    TRY(field_ptr(h, field, &d, &es));
    const uint32_t N = (uint32_t)h->L.N;
    if (!host || row_begin > N || n_rows > N - row_begin) return eq_fail(EQ_ERR_INVALID, "bad row range");
    if (h->world > 1 && n_rows > 0 && ((int)row_begin < h->L.row0 || (int)(row_begin + n_rows) > h->L.row1))
        return eq_fail(EQ_ERR_INVALID, "rows [%u,%u) are not all owned by rank %d (owns [%d,%d))", row_begin,
                       row_begin + n_rows, h->rank, h->L.row0, h->L.row1);
    if (n_rows > 0) {
        ProfScope ps(h, CAT_OTHER, 0);
        CU(cudaMemcpy2DAsync(host, (size_t)N * es, static_cast<char *>(d) + (size_t)row_begin * h->L.P * es,
                             (size_t)h->L.P * es, (size_t)N * es, n_rows, cudaMemcpyDeviceToHost, h->stream));
    }
    return check_device_error(h);   // synchronises the stream
}

int eq_upload(eq_fluid *h, int field, const void *host, size_t bytes) {
    NEED(h);
    const size_t es = field == EQ_F_CELLS ? 1 : sizeof(float);
    if (bytes != (size_t)h->L.N * h->L.N * es) return eq_fail(EQ_ERR_INVALID, "eq_upload: expected %zu bytes, got %zu", (size_t)h->L.N * h->L.N * es, bytes);
    return eq_upload_rows(h, field, 0, (uint32_t)h->L.N, host);
}

int eq_download(eq_fluid *h, int field, void *host, size_t bytes) {
    NEED(h);
    if (h->world > 1) return eq_fail(EQ_ERR_STATE, "eq_download of a row-slab handle: use eq_owned_rows + eq_download_rows");
    const size_t es = field == EQ_F_CELLS ? 1 : sizeof(float);
    if (bytes != (size_t)h->L.N * h->L.N * es) return eq_fail(EQ_ERR_INVALID, "eq_download: expected %zu bytes, got %zu", (size_t)h->L.N * h->L.N * es, bytes);
    return eq_download_rows(h, field, 0, (uint32_t)h->L.N, host);
}

int eq_owned_rows(eq_fluid *h, uint32_t *row_begin, uint32_t *row_end) {
    if (!h || !row_begin || !row_end) return eq_fail(EQ_ERR_INVALID, "null argument");
    *row_begin = (uint32_t)h->L.row0;
    *row_end = (uint32_t)h->L.row1;
    return EQ_OK;
}

static int f32_field(eq_fluid *h, int id, float **out) {
    if (id < 0 || id >= 6) return eq_fail(EQ_ERR_INVALID, "field id %d is not an f32 field", id);
    *out = h->f[id];
    return EQ_OK;
}

static bool valid_orient(int o) { return o == EQ_ADJUST_ROW || o == EQ_ADJUST_COLUMN || o == EQ_PASSIVE; }

int eq_op_set_boundaries(eq_fluid *h, int orientation, int field) {
    NEED(h);
    float *x;
    TRY(f32_field(h, field, &x));
    if (!valid_orient(orientation)) return eq_fail(EQ_ERR_INVALID, "bad orientation");
    return set_boundaries(h, orientation, x);
}

// dense source field: x += scale * s on the whole grid (Stam's add_source; the reference only has point sources)
int eq_op_add_source(eq_fluid *h, int x_field, int s_field, float scale) {
    NEED(h);
    float *x, *s;
    TRY(f32_field(h, x_field, &x));
    TRY(f32_field(h, s_field, &s));
    if (x == s) return eq_fail(EQ_ERR_INVALID, "add_source needs two different fields");
    const int quads = (h->L.N + 3) / 4;
    dim3 grid((quads + 255) / 256, std::min(h->L.N, 148 * 8));
    ProfScope ps(h, CAT_OTHER, 1);
    EQ_LAUNCH(k_add_field, grid, 256, 0, h->stream, x, s, scale, h->L);
    return check_launch("k_add_field");
}

int eq_op_lin_solve(eq_fluid *h, int orientation, int x_field, int x0_field, float a, float c, int64_t iters) {
    NEED(h);
    float *x, *x0;
    TRY(f32_field(h, x_field, &x));
    TRY(f32_field(h, x0_field, &x0));
    if (!valid_orient(orientation) || x == x0 || iters < 0) return eq_fail(EQ_ERR_INVALID, "bad lin_solve arguments");
    LinSolveReq r{orientation, x, x0, a, c};
    return lin_solve(h, &r, 1, iters);
}

int eq_op_diffuse(eq_fluid *h, int orientation, int x_field, int x0_field, float diffusion, int64_t iters) {
    NEED(h);
    float *x, *x0;
    TRY(f32_field(h, x_field, &x));
    TRY(f32_field(h, x0_field, &x0));
    if (!valid_orient(orientation) || x == x0 || iters < 0) return eq_fail(EQ_ERR_INVALID, "bad diffuse arguments");
    LinSolveReq r = diffuse_req(h, orientation, x, x0, diffusion);
    return lin_solve(h, &r, 1, iters);
}

int eq_op_project(eq_fluid *h, int vx_field, int vy_field, int p_field, int div_field, int64_t iters) {
    NEED(h);
    float *vx, *vy, *p, *div;
    TRY(f32_field(h, vx_field, &vx));
    TRY(f32_field(h, vy_field, &vy));
    TRY(f32_field(h, p_field, &p));
    TRY(f32_field(h, div_field, &div));
    if (vx == vy || vx == p || vx == div || vy == p || vy == div || p == div || iters < 0)
        return eq_fail(EQ_ERR_INVALID, "project needs four distinct fields");
    return project(h, vx, vy, p, div, iters);
}

int eq_op_advect(eq_fluid *h, int orientation, int d_field, int d0_field, int vx_field, int vy_field) {
    NEED(h);
    float *d, *d0, *vx, *vy;
    TRY(f32_field(h, d_field, &d));
    TRY(f32_field(h, d0_field, &d0));
    TRY(f32_field(h, vx_field, &vx));
    TRY(f32_field(h, vy_field, &vy));
    if (!valid_orient(orientation) || d == d0 || d == vx || d == vy) return eq_fail(EQ_ERR_INVALID, "bad advect arguments");
    return advect(h, orientation, d, d0, 0, nullptr, nullptr, vx, vy);
}

// ---------------------------------------------------------------------------
// per-frame snapshots and the colour map
// ---------------------------------------------------------------------------
static uint32_t pack_rgba(const uint8_t c[4]) {
    return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
}

int eq_snapshot_begin(eq_fluid *h, int kind, int slot, const EqColors *colors, void *host_dst, size_t bytes) {
    NEED(h);
    if (slot < 0 || slot >= EQ_SNAPSHOT_SLOTS) return eq_fail(EQ_ERR_INVALID, "snapshot slot %d not in [0,%d)", slot, EQ_SNAPSHOT_SLOTS);
    if (kind != EQ_SNAP_DENSITY && kind != EQ_SNAP_RGBA) return eq_fail(EQ_ERR_INVALID, "unknown snapshot kind %d", kind);
    if (kind == EQ_SNAP_RGBA && !colors) return eq_fail(EQ_ERR_INVALID, "EQ_SNAP_RGBA needs colours");
    const EqLayout L = h->L;
    const size_t rows = (size_t)(L.row1 - L.row0), want = rows * L.N * 4;
    if (!host_dst || bytes != want) return eq_fail(EQ_ERR_INVALID, "eq_snapshot_begin: expected %zu bytes (%zu owned rows), got %zu", want, rows, bytes);
    if (!h->copy_stream) CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->snap_buf[0]) {   // all slots at once: an allocation in the middle of a run would synchronise the device
        for (int i = 0; i < EQ_SNAPSHOT_SLOTS; ++i) {
            CU(cudaMalloc(&h->snap_buf[i], want));
            CU(cudaEventCreate(&h->snap_ready[i]));
            CU(cudaEventCreate(&h->snap_done[i]));
        }
    }
    // the staging slot may still be on its way to the host from its previous use
    if (h->snap_used[slot]) CU(cudaStreamWaitEvent(h->stream, h->snap_done[slot], 0));
    {
        ProfScope ps(h, CAT_OTHER, kind == EQ_SNAP_RGBA ? 1 : 0);
        if (kind == EQ_SNAP_DENSITY) {
            CU(cudaMemcpy2DAsync(h->snap_buf[slot], (size_t)L.N * 4, h->f[EQ_F_DENSITY] + (size_t)L.row0 * L.P, (size_t)L.P * 4,
                                 (size_t)L.N * 4, rows, cudaMemcpyDeviceToDevice, h->stream));
        } else {
            EQ_LAUNCH(k_render_rgba, row_grid(h, (int)rows), 256, 0, h->stream, h->f[EQ_F_DENSITY], h->cells,
                      static_cast<uint32_t *>(h->snap_buf[slot]), pack_rgba(colors->world), pack_rgba(colors->fluid),
                      pack_rgba(colors->obstacle), L);
            TRY(check_launch("k_render_rgba"));
        }
    }
    CU(cudaEventRecord(h->snap_ready[slot], h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->snap_ready[slot], 0));
    CU(cudaMemcpyAsync(host_dst, h->snap_buf[slot], want, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(cudaEventRecord(h->snap_done[slot], h->copy_stream));
    h->snap_used[slot] = true;
    return EQ_OK;
}

int eq_snapshot_wait(eq_fluid *h, int slot) {
    NEED(h);
    if (slot < 0 || slot >= EQ_SNAPSHOT_SLOTS) return eq_fail(EQ_ERR_INVALID, "snapshot slot %d not in [0,%d)", slot, EQ_SNAPSHOT_SLOTS);
    if (!h->snap_used[slot]) return eq_fail(EQ_ERR_STATE, "no snapshot was started in slot %d", slot);
    CU(cudaEventSynchronize(h->snap_done[slot]));
    return EQ_OK;
}

int eq_render_rgba(eq_fluid *h, const EqColors *colors, void *host_rgba, size_t bytes) {
    TRY(eq_snapshot_begin(h, EQ_SNAP_RGBA, 0, colors, host_rgba, bytes));
    return eq_snapshot_wait(h, 0);
}

int eq_divergence_l2(eq_fluid *h, int vx_field, int vy_field, double *out) {
    NEED(h);
    float *vx, *vy;
    TRY(f32_field(h, vx_field, &vx));
    TRY(f32_field(h, vy_field, &vy));
    if (!out) return eq_fail(EQ_ERR_INVALID, "null out");
    double *d = nullptr;
    CU(cudaMalloc(&d, sizeof(double)));
    CU(cudaMemsetAsync(d, 0, sizeof(double), h->stream));
    TRY(halo_xchg(h, vy));
    EQ_LAUNCH(k_divergence_sq, row_grid(h, owned_interior_rows(h)), 256, 0, h->stream, vx, vy, d, h->L);
    int r = check_launch("k_divergence_sq");
    double v = 0.0;
    if (r == EQ_OK && cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) r = EQ_ERR_CUDA;
    if (r == EQ_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) r = EQ_ERR_CUDA;
    cudaFree(d);
    if (r != EQ_OK) return eq_fail(r, "eq_divergence_l2 failed");
    *out = sqrt(v);
    return EQ_OK;
}

int eq_set_stream(eq_fluid *h, void *cuda_stream) {
    NEED(h);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return EQ_OK;
}

int eq_timer_start(eq_fluid *h) {
    NEED(h);
    CU(cudaEventRecord(h->ev0, h->stream));
    return EQ_OK;
}

int eq_timer_stop(eq_fluid *h, float *ms) {
    NEED(h);
    if (!ms) return eq_fail(EQ_ERR_INVALID, "null out");
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return EQ_OK;
}

int eq_profile_enable(eq_fluid *h, int on) {
    NEED(h);
    TRY(prof_collect(h));
    h->prof_on = on != 0;
    return EQ_OK;
}

int eq_profile_reset(eq_fluid *h) {
    NEED(h);
    TRY(prof_collect(h));
    memset(h->prof_ms, 0, sizeof(h->prof_ms));
    memset(h->prof_launches, 0, sizeof(h->prof_launches));
    h->prof_cell_iters = 0;
    h->prof_steps = 0;
    return EQ_OK;
}

int eq_profile_get(eq_fluid *h, EqProfile *out) {
    NEED(h);
    if (!out) return eq_fail(EQ_ERR_INVALID, "null out");
    TRY(prof_collect(h));
    out->lin_solve_ms = h->prof_ms[CAT_LS];
    out->lin_solve_launches = h->prof_launches[CAT_LS];
    out->lin_solve_cell_iters = h->prof_cell_iters;
    out->advect_ms = h->prof_ms[CAT_ADV];
    out->advect_launches = h->prof_launches[CAT_ADV];
    out->project_ms = h->prof_ms[CAT_PROJ];
    out->project_launches = h->prof_launches[CAT_PROJ];
    out->boundary_ms = h->prof_ms[CAT_BND];
    out->boundary_launches = h->prof_launches[CAT_BND];
    out->other_ms = h->prof_ms[CAT_OTHER];
    out->other_launches = h->prof_launches[CAT_OTHER];
    out->steps = h->prof_steps;
    return EQ_OK;
}

int eq_host_alloc(void **out, size_t bytes) {
    if (!out) return eq_fail(EQ_ERR_INVALID, "null out");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return EQ_OK;
}

int eq_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return EQ_OK;
}

int eq_l2_flush(eq_fluid *h) {
    NEED(h);
    if (!h->l2buf) {
        h->l2bytes = (size_t)256 << 20;   // > 126 MB of L2
        CU(cudaMalloc(&h->l2buf, h->l2bytes));
    }
    CU(cudaMemsetAsync(h->l2buf, 0, h->l2bytes, h->stream));
    return EQ_OK;
}

// ---- multi-GPU rendezvous ---------------------------------------------------------------
// Each rank exports one blob; the launcher gathers them (torch.distributed / any side channel)
// and hands every rank the concatenation.  Peers in other processes are mapped with CUDA IPC,
// peers in the same process (several handles, one per device) with plain peer access.
#define EQ_IPC_ITEMS 11
struct EqIpcBlob {
    uint32_t magic, rank, world, size;
    int32_t device, pad;
    int64_t pid;
    uint64_t ptr[EQ_IPC_ITEMS];
    cudaIpcMemHandle_t handle[EQ_IPC_ITEMS];
};

static void ipc_items(eq_fluid *h, void *items[EQ_IPC_ITEMS]) {
    for (int i = 0; i < 6; ++i) items[i] = h->f[i];
    items[6] = h->raw[0];
    items[7] = h->raw[1];
    items[8] = h->flags;
    items[9] = h->sync;
    items[10] = h->rb_tmp;
}

int eq_ipc_blob_bytes(void) { return (int)sizeof(EqIpcBlob); }

int eq_ipc_export(eq_fluid *h, void *blob, size_t capacity) {
    NEED(h);
    if (!blob || capacity < sizeof(EqIpcBlob)) return eq_fail(EQ_ERR_INVALID, "blob buffer too small (%zu needed)", sizeof(EqIpcBlob));
    CU(cudaStreamSynchronize(h->stream));   // allocations are zeroed before anybody maps them
    EqIpcBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = 0x45514950u;
    b.rank = (uint32_t)h->rank;
    b.world = (uint32_t)h->world;
    b.size = h->prm.size;
    b.device = h->dev;
    b.pid = (int64_t)getpid();
    void *items[EQ_IPC_ITEMS];
    ipc_items(h, items);
    for (int i = 0; i < EQ_IPC_ITEMS; ++i) {
        b.ptr[i] = (uint64_t)(uintptr_t)items[i];
        CU(cudaIpcGetMemHandle(&b.handle[i], items[i]));
    }
    memcpy(blob, &b, sizeof(b));
    return EQ_OK;
}

int eq_ipc_attach(eq_fluid *h, const void *blobs, size_t blob_bytes, int world) {
    NEED(h);
    if (!blobs || blob_bytes != sizeof(EqIpcBlob) || world != h->world)
        return eq_fail(EQ_ERR_INVALID, "eq_ipc_attach: need %d blobs of %zu bytes", h->world, sizeof(EqIpcBlob));
    for (int r = 0; r < world; ++r) {
        EqIpcBlob b;
        memcpy(&b, static_cast<const char *>(blobs) + (size_t)r * blob_bytes, sizeof(b));
        if (b.magic != 0x45514950u || (int)b.rank != r || (int)b.world != world || b.size != h->prm.size)
            return eq_fail(EQ_ERR_COMM, "blob %d does not describe rank %d of %d for a %u grid", r, r, world, h->prm.size);
        if (r == h->rank) continue;
        void *mapped[EQ_IPC_ITEMS];
        if (b.pid == (int64_t)getpid()) {
            if (b.device != h->dev) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return eq_fail(EQ_ERR_COMM, "cudaDeviceEnablePeerAccess(%d) failed: %s", b.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            for (int i = 0; i < EQ_IPC_ITEMS; ++i) mapped[i] = (void *)(uintptr_t)b.ptr[i];
        } else {
            for (int i = 0; i < EQ_IPC_ITEMS; ++i) {
                cudaError_t e = cudaIpcOpenMemHandle(&mapped[i], b.handle[i], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) return eq_fail(EQ_ERR_COMM, "cudaIpcOpenMemHandle (rank %d, item %d) failed: %s", r, i, cudaGetErrorString(e));
            }
            h->peer_ipc[r] = true;
        }
        for (int i = 0; i < 6; ++i) h->peer_f[r][i] = static_cast<float *>(mapped[i]);
        h->peer_raw[r][0] = static_cast<float *>(mapped[6]);
        h->peer_raw[r][1] = static_cast<float *>(mapped[7]);
        h->peer_flags[r] = static_cast<unsigned *>(mapped[8]);
        h->peer_sync[r] = static_cast<unsigned *>(mapped[9]);
        h->peer_tmp[r] = static_cast<float *>(mapped[10]);
    }
    h->attached = true;
    return EQ_OK;
}

}  // extern "C"
