/*
 * oracle/fluid_ref.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Scalar, single-threaded C restatement of the stable-fluids solver in the
 * reference's src/simulation/fluid.rs. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library,
 * and only as the checker / CPU baseline -- never as the product path.
 *
 * PARITY PINNING: the reference's own tests pin only (a) the row-major index
 * layout (fluid.rs:626-635), (b) the 1408 wall cells of the default scene
 * (renderer_helpers.rs:222-252) and (c) Rectangle validation
 * (obstacle.rs:100-115). All three are checked in tests/test_oracle.py.  No
 * reference test pins any value produced by step(); the reference is Rust and
 * no Rust toolchain exists in the build image, so the numerical behaviour of
 * step() is "PARITY UNPINNED" beyond the source text itself.  Mitigation: a
 * second, independently written restatement (oracle/pyref.py) must agree with
 * this one bit-for-bit (tests/test_oracle.py).
 *
 * Build flags that matter: -O2 -ffp-contract=off -fno-fast-math (Rust never
 * fuses mul+add and never reassociates).
 */
#ifndef EQ_ORACLE_FLUID_REF_H
#define EQ_ORACLE_FLUID_REF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fluid.rs:21-29 `enum Orientation` */
enum { REF_ADJUST_ROW = 0, REF_ADJUST_COLUMN = 1, REF_PASSIVE = 2 };

/* cells_type encoding used across this repo: 0 = NoWall, 1 = DefaultWall
 * (fluid.rs:11-17). */
enum { REF_NO_WALL = 0, REF_DEFAULT_WALL = 1 };

/* field ids, shared with include/equilibrium_cuda.h */
enum {
    REF_F_DENSITY = 0, REF_F_VX = 1, REF_F_VY = 2,
    REF_F_VX0 = 3, REF_F_VY0 = 4, REF_F_SCRATCH = 5, REF_F_CELLS = 6
};

/* fluid.rs:51-81 `struct Fluid`.  `rows` == `size` for every reference
 * scene; rows < size exists only so bench.py can time a row band of a very
 * wide grid (coefficients keep using `size`, loops over j use `rows`). */
typedef struct ref_fluid {
    uint32_t size;
    uint32_t rows;
    float delta_t;        /* configs.rs:7  */
    int64_t frames;       /* configs.rs:9  (also the GS iteration count, quirk Q1) */
    int64_t gs_iterations;/* 0 => use `frames` exactly like fluid.rs:445 */
    float diffusion;      /* configs.rs:40 */
    float viscosity;      /* configs.rs:42 (`viscousity`) */
    float *scratch_space, *density, *velocities_x, *velocities_y,
          *velocities_x0, *velocities_y0;
    uint8_t *cells_type;
} ref_fluid;

/* Fluid::new (fluid.rs:93-110): allocate zeroed fields, NoWall mask, init(). */
ref_fluid *ref_fluid_new(uint32_t size, uint32_t rows, float delta_t, int64_t frames,
                         int64_t gs_iterations, float diffusion, float viscosity);
void ref_fluid_free(ref_fluid *f);
/* Fluid::init (fluid.rs:602-606); Default (fluid.rs:83-89) = new + init again. */
void ref_fluid_init(ref_fluid *f);
/* fluid.rs:120-131 */
void ref_add_density(ref_fluid *f, uint32_t x, uint32_t y, float amount);
void ref_add_velocity(ref_fluid *f, uint32_t x, uint32_t y, float ax, float ay);
/* fluid.rs:610-619 with the two approximate points of a Rectangle */
void ref_fill_rect(ref_fluid *f, int64_t x0, int64_t y0, int64_t x1, int64_t y1);
/* obstacle.rs:74-87 Rectangle::are_all_points_valid (1 = valid) */
int ref_rect_valid(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t size);
/* fluid.rs:437-524 */
void ref_fluid_step(ref_fluid *f);
/* raw pointer to one of the seven arrays (REF_F_*) */
void *ref_fluid_field(ref_fluid *f, int field_id);

/* The building blocks, exported so each CUDA kernel can be checked alone. */
void ref_set_boundaries(int orientation, float *x, uint32_t size, uint32_t rows,
                        const uint8_t *cells);                       /* fluid.rs:252-272 */
void ref_lin_solve(int orientation, float *x, const float *x0, float a, float c,
                   uint32_t size, uint32_t rows, int64_t iters,
                   const uint8_t *cells);                            /* fluid.rs:301-325 */
void ref_diffuse(int orientation, float *x, const float *x0, float diffusion,
                 uint32_t size, uint32_t rows, float delta_t, int64_t iters,
                 const uint8_t *cells);                              /* fluid.rs:276-298 */
void ref_project(float *vx, float *vy, float *p, float *div, uint32_t size,
                 uint32_t rows, int64_t iters, const uint8_t *cells); /* fluid.rs:330-375 */
void ref_advect(int orientation, float *d, const float *d0, const float *vx,
                const float *vy, uint32_t size, uint32_t rows, float delta_t,
                const uint8_t *cells);                               /* fluid.rs:378-432 */

/* Red-black variant of lin_solve used ONLY to state the tolerance of the
 * product's red-black fast path: same formula and iteration count as
 * fluid.rs:301-325, but each iteration updates cells with (i+j) even first,
 * then (i+j) odd, then set_boundaries. */
void ref_lin_solve_red_black(int orientation, float *x, const float *x0, float a, float c,
                             uint32_t size, uint32_t rows, int64_t iters,
                             const uint8_t *cells);

/* The pixel loop of RenderingListener::render_image (renderer_helpers.rs:145-167): one RGBA
 * pixel per cell from density + cells_type.  colours = world, fluid, obstacle as r,g,b,a bytes
 * (renderer_helpers.rs:122-143).  out = size*rows*4 bytes. */
void ref_render_rgba(const float *density, const uint8_t *cells, uint32_t size, uint32_t rows,
                     const uint8_t world[4], const uint8_t fluid[4], const uint8_t obstacle[4],
                     uint8_t *out);

/* Device-side add_noise and dense sources (structure of fluid.rs:575-599). */
void ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void ref_noise_impulse(uint64_t seed, uint64_t frame, uint32_t size, float cos_t, float sin_t, float gain,
                       uint32_t xy[2], float a[2]);
void ref_add_source(float *x, const float *s, float scale, size_t cells);

#ifdef __cplusplus
}
#endif
#endif
