/*
 * include/equilibrium_cuda.h -- C ABI of libequilibrium_cuda.so
 *
 * B200 (sm_100a) implementation of the per-frame stable-fluids step of
 * vkabadzhova/equilibrium.  The reference has no FFI of its own: its "operator
 * API" is the inherent-method surface of `Fluid` (src/simulation/fluid.rs).
 * Every entry point below names the reference item it replaces (paths are
 * relative to the reference repo).  INTEGRATION.md shows the Rust `-sys`
 * binding and the feature-gated `Fluid` that calls it.
 *
 * Conventions
 *  - plain C, opaque handle, no exceptions; every call returns EQ_OK (0) or a
 *    negative EqStatus; eq_last_error() gives a thread-local message.
 *  - fields are f32, row-major, idx = x + y*size (fluid.rs:31-35, test
 *    fluid.rs:626-635); cells_type is u8: 0 = NoWall, 1 = DefaultWall
 *    (fluid.rs:11-17; note Rust's discriminants are the other way round).
 *  - the handle owns device memory; host pointers are borrowed for the call.
 *  - a handle may be used from one thread at a time (the reference moves its
 *    Fluid to the simulation thread, renderer.rs:125-128); calls are
 *    stream-ordered on the handle's stream; eq_download/eq_sync wait.
 *  - there is NO CPU fallback: without a CUDA device eq_create fails.
 */
#ifndef EQUILIBRIUM_CUDA_H
#define EQUILIBRIUM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQ_ABI_VERSION 1

typedef struct eq_fluid eq_fluid;

typedef enum EqStatus {
    EQ_OK = 0,
    EQ_ERR_INVALID = -1,   /* bad argument (the reference would panic) */
    EQ_ERR_CUDA = -2,      /* CUDA runtime error, see eq_last_error() */
    EQ_ERR_NOMEM = -3,
    EQ_ERR_STATE = -4,     /* e.g. setter used while a step_n is being captured */
    EQ_ERR_TIMEOUT = -5,   /* wavefront watchdog fired (never expected) */
    EQ_ERR_COMM = -6       /* halo-exchange transport error */
} EqStatus;

/* Gauss-Seidel ordering of lin_solve (fluid.rs:301-325). */
typedef enum EqMode {
    EQ_MODE_EXACT = 0,     /* wavefront kernel reproducing the reference's lexicographic
                              in-place order: results bit-identical to fluid.rs */
    EQ_MODE_RED_BLACK = 1  /* red-black ordering, same iteration count; tolerance-checked */
} EqMode;

/* fluid.rs:21-29 `enum Orientation` */
typedef enum EqOrientation { EQ_ADJUST_ROW = 0, EQ_ADJUST_COLUMN = 1, EQ_PASSIVE = 2 } EqOrientation;

/* the seven arrays of `struct Fluid` (fluid.rs:51-81) */
typedef enum EqField {
    EQ_F_DENSITY = 0,  /* pub density          fluid.rs:61 */
    EQ_F_VX = 1,       /* pub velocities_x     fluid.rs:65 */
    EQ_F_VY = 2,       /* pub velocities_y     fluid.rs:69 */
    EQ_F_VX0 = 3,      /* velocities_x0        fluid.rs:73 */
    EQ_F_VY0 = 4,      /* velocities_y0        fluid.rs:77 */
    EQ_F_SCRATCH = 5,  /* scratch_space        fluid.rs:59 */
    EQ_F_CELLS = 6     /* pub cells_type (u8)  fluid.rs:80 */
} EqField;

/* SimulationConfigs (configs.rs:5-12) + the numeric part of FluidConfigs
 * (configs.rs:37-48) + what the CUDA path adds. */
typedef struct EqParams {
    uint32_t size;          /* SimulationConfigs::size, 20 <= size <= 32768 */
    float delta_t;          /* SimulationConfigs::delta_t */
    int64_t frames;         /* SimulationConfigs::frames (frame count AND, in the
                               reference, the GS iteration count: fluid.rs:445) */
    int64_t gs_iterations;  /* 0 => use `frames` like the reference does */
    float diffusion;        /* FluidConfigs::diffusion */
    float viscosity;        /* FluidConfigs::viscousity */
    int32_t mode;           /* EqMode */
    int32_t device;         /* CUDA device ordinal */
    /* Row-slab decomposition over several GPUs, one process per GPU
     * (SURVEY.md 8e).  world <= 1: single GPU, the other fields are ignored. */
    int32_t rank;
    int32_t world;
    uint8_t comm_id[128];   /* reserved */
} EqParams;

/* Point source applied before the step of frame `frame` (what add_noise does,
 * fluid.rs:575-599 -> add_velocity; plus add_density fluid.rs:120-124). */
typedef struct EqSource {
    int64_t frame;
    uint32_t x, y;
    float d_vx, d_vy, d_density;
} EqSource;

/* Device-side add_noise (fluid.rs:575-599).  Frame f of a call
 * draws (rx, ry) in [0,size)^2 from Philox4x32-10(counter = first_frame + f,
 * key = seed), rotates that point about the centre (size/2, size/2) with
 * (cos_t, sin_t) and adds gain * rotated point to the centre cell's velocity
 * (add_velocity, fluid.rs:127-131).  The reference's angle is a constant of
 * delta_t (fluid.rs:578-583) and its gain is 2.0 (fluid.rs:595-596); its RNG is
 * an unseeded thread_rng, so no sequence of it can be reproduced. */
typedef struct EqNoise {
    uint64_t seed;
    uint64_t first_frame;
    float cos_t, sin_t;
    float gain;
    float reserved;
} EqNoise;

/* Device-side time of each phase of the steps run since eq_profile_reset,
 * measured with CUDA events on the handle's stream (enable with
 * eq_profile_enable).  Times in ms, launches = kernel launches counted. */
typedef struct EqProfile {
    double lin_solve_ms; int64_t lin_solve_launches; int64_t lin_solve_cell_iters;
    double advect_ms;    int64_t advect_launches;
    double project_ms;   int64_t project_launches;   /* divergence + gradient kernels */
    double boundary_ms;  int64_t boundary_launches;  /* stand-alone set_boundaries passes */
    double other_ms;     int64_t other_launches;     /* sources, copies, halo exchange */
    int64_t steps;
} EqProfile;

const char *eq_last_error(void);
int eq_abi_version(void);
/* number of visible CUDA devices (0 => nothing will work; no CPU fallback) */
int eq_device_count(void);

/* Fluid::new (fluid.rs:93-110): zeroed fields, NoWall mask, then init()
 * (fluid.rs:602-606: v=(1,1), 21x21 density block of 0.9, frame walls). */
int eq_create(const EqParams *params, eq_fluid **out);
/* Drop */
int eq_destroy(eq_fluid *h);
/* #[derive(Clone)] (fluid.rs:51): deep copy, used per frame by the caller
 * (renderer_helpers.rs:61-65). */
int eq_clone(eq_fluid *h, eq_fluid **out);
/* Fluid::init (fluid.rs:602-606).  Default::default() = eq_create + eq_init_default
 * (fluid.rs:83-89 initialises twice). */
int eq_init_default(eq_fluid *h);

/* add_density (fluid.rs:120-124) / add_velocity (fluid.rs:127-131); x,y clamped like idx!. */
int eq_add_density(eq_fluid *h, uint32_t x, uint32_t y, float amount);
int eq_add_velocity(eq_fluid *h, uint32_t x, uint32_t y, float amount_x, float amount_y);

/* Rectangle::are_all_points_valid (obstacle.rs:74-87): 1 valid, 0 invalid. */
int eq_rect_valid(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t size);
/* fill_obstacle (fluid.rs:610-619) for a Rectangle's two approximate points:
 * half-open [x0,x1) x [y0,y1), coordinates clamped like idx!. */
int eq_fill_rect(eq_fluid *h, int64_t x0, int64_t y0, int64_t x1, int64_t y1);
/* "remove obstacles": back to the frame-only mask of init_walls (fluid.rs:552-570);
 * the reference does this by building a new Fluid (renderer.rs:145-149). */
int eq_reset_walls(eq_fluid *h);

/* Parameter setters: legal between steps only.  size/device/rank/world are fixed. */
int eq_set_params(eq_fluid *h, const EqParams *params);
int eq_get_params(eq_fluid *h, EqParams *out);

/* Fluid::step (fluid.rs:437-524), enqueued on the handle's stream. */
int eq_step(eq_fluid *h);
/* n frames without host synchronisation; `sources` (may be NULL) must be sorted
 * by frame, frames counted from 0 for this call
 * (CurrentSimulation::simulate's loop, renderer_helpers.rs:54-66). */
int eq_step_n(eq_fluid *h, int64_t n, const EqSource *sources, int64_t n_sources);
/* add_noise (fluid.rs:575-599) on the device: the impulse of counter noise->first_frame, no step. */
int eq_add_noise(eq_fluid *h, const EqNoise *noise);
/* n x { device-side add_noise; step() }: no source record crosses the bus. */
int eq_step_n_noise(eq_fluid *h, int64_t n, const EqNoise *noise);
/* wait for everything enqueued; returns EQ_ERR_TIMEOUT if a wavefront watchdog fired */
int eq_sync(eq_fluid *h);

/* Raw field access (the pub fields of Fluid, plus the private ones for
 * checkpoint/parity): `bytes` must be size*size*4 (f32 fields) or size*size
 * (EQ_F_CELLS).  The frame of an uploaded mask must be all DefaultWall
 * (init_walls always marks it and the reference has no way to clear it). */
int eq_upload(eq_fluid *h, int field, const void *host, size_t bytes);
int eq_download(eq_fluid *h, int field, void *host, size_t bytes);
/* Multi-GPU: rows [row_begin,row_begin+n_rows) must lie in the slab this rank owns. */
int eq_upload_rows(eq_fluid *h, int field, uint32_t row_begin, uint32_t n_rows, const void *host);
int eq_download_rows(eq_fluid *h, int field, uint32_t row_begin, uint32_t n_rows, void *host);
/* rows owned by this rank: [*row_begin, *row_end) */
int eq_owned_rows(eq_fluid *h, uint32_t *row_begin, uint32_t *row_end);

/* The building blocks of step(), each on the handle's own arrays, so that every
 * kernel can be checked against the oracle in isolation. */
int eq_op_set_boundaries(eq_fluid *h, int orientation, int field);                 /* fluid.rs:252-272 */
int eq_op_lin_solve(eq_fluid *h, int orientation, int x_field, int x0_field, float a, float c,
                    int64_t iters);                                                /* fluid.rs:301-325 */
int eq_op_diffuse(eq_fluid *h, int orientation, int x_field, int x0_field, float diffusion,
                  int64_t iters);                                                  /* fluid.rs:276-298 */
int eq_op_project(eq_fluid *h, int vx_field, int vy_field, int p_field, int div_field,
                  int64_t iters);                                                  /* fluid.rs:330-375 */
int eq_op_advect(eq_fluid *h, int orientation, int d_field, int d0_field, int vx_field,
                 int vy_field);                                                    /* fluid.rs:378-432 */
/* Dense source field: x += scale * s on every cell, as Stam's add_source does; the
 * reference itself only has the point sources above (fluid.rs:120-131). */
int eq_op_add_source(eq_fluid *h, int x_field, int s_field, float scale);

/* ---- the caller's side of the frame loop -----------------------
 * After every step() the reference deep-copies the whole Fluid and sends it to the render
 * thread (`fluid.clone()` + mpsc, renderer_helpers.rs:61-65), which turns density + cells_type
 * into RGBA pixels (render_image, renderer_helpers.rs:115-167) before the JPEG encode.  Here the
 * frame leaves the GPU as ONE array -- f32 density or the finished RGBA pixels -- through a
 * pinned-memory snapshot that overlaps the next step(). */
typedef struct EqColors {      /* Color32 r,g,b,a of FluidConfigs::world_color / fluid_color  */
    uint8_t world[4];          /* (configs.rs:37-48) and RenderingListener::obstacles_color   */
    uint8_t fluid[4];          /* (renderer_helpers.rs:86-92)                                 */
    uint8_t obstacle[4];
} EqColors;
typedef enum EqSnapshotKind {
    EQ_SNAP_DENSITY = 0,       /* f32 density, row-major size*size (the rows this rank owns)  */
    EQ_SNAP_RGBA = 1           /* 4 x u8 per cell, the pixel rule of renderer_helpers.rs:145-167 */
} EqSnapshotKind;
#define EQ_SNAPSHOT_SLOTS 2
/* Enqueue a snapshot of the CURRENT state (stream-ordered after the steps issued so far): a
 * device-side copy / colour-map into staging slot `slot`, then an asynchronous device->host copy
 * on a second stream into `host_dst` (pinned memory from eq_host_alloc for real overlap;
 * `bytes` = owned_rows * size * 4).  The call returns at once; later steps may run while the
 * copy is in flight.  Re-using a slot waits (on the device) for its previous copy. */
int eq_snapshot_begin(eq_fluid *h, int kind, int slot, const EqColors *colors, void *host_dst, size_t bytes);
/* Block until the snapshot in `slot` has landed in its host buffer. */
int eq_snapshot_wait(eq_fluid *h, int slot);
/* Synchronous convenience: render_image's pixel loop for the owned rows into host memory. */
int eq_render_rgba(eq_fluid *h, const EqColors *colors, void *host_rgba, size_t bytes);

/* Diagnostics for BASELINE config 5: L2 norm of the velocity divergence
 * (same stencil as fluid.rs:341-345) over the owned interior. */
int eq_divergence_l2(eq_fluid *h, int vx_field, int vy_field, double *out);

/* Streams, timing, pinned host memory. */
int eq_set_stream(eq_fluid *h, void *cuda_stream);   /* NULL => the handle's own stream */
int eq_timer_start(eq_fluid *h);                     /* cudaEventRecord on the stream */
int eq_timer_stop(eq_fluid *h, float *ms);           /* record + synchronize + elapsed */
int eq_profile_enable(eq_fluid *h, int on);
int eq_profile_reset(eq_fluid *h);
int eq_profile_get(eq_fluid *h, EqProfile *out);
int eq_host_alloc(void **out, size_t bytes);         /* cudaHostAlloc (pinned) */
int eq_host_free(void *p);
int eq_l2_flush(eq_fluid *h);                        /* overwrite a buffer larger than L2 */

/* Multi-GPU rendezvous (row slabs, one process per GPU): after eq_create with world > 1 every
 * rank exports a blob of eq_ipc_blob_bytes() bytes, the launcher gathers them
 * (torch.distributed / any side channel) and every rank attaches the rank-ordered
 * concatenation.  Neighbours are then read and written directly over NVLink by the kernels. */
int eq_ipc_blob_bytes(void);
int eq_ipc_export(eq_fluid *h, void *blob, size_t capacity);
int eq_ipc_attach(eq_fluid *h, const void *blobs, size_t blob_bytes, int world);

#ifdef __cplusplus
}
#endif
#endif /* EQUILIBRIUM_CUDA_H */
