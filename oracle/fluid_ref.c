/*
 * oracle/fluid_ref.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See fluid_ref.h for the scope and the "parity unpinned" statement.
 *
 * Every function names the reference lines (relative to /root/reference/) it
 * restates.  Loop order, expression trees (left-to-right adds, no FMA) and the
 * naive full-grid set_boundaries are kept exactly as in the reference: this
 * file is the specification the CUDA path is compared against, bit for bit.
 */
#include "fluid_ref.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* fluid.rs:31-35  idx!(x, y, size) = x.clamp(0,size-1) + y.clamp(0,size-1)*size.
 * `rows` bounds y; rows == size in every reference scene. */
static inline int64_t clampi(int64_t v, int64_t lo, int64_t hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}
static inline size_t IDX(int64_t x, int64_t y, int64_t size, int64_t rows) {
    return (size_t)(clampi(x, 0, size - 1) + clampi(y, 0, rows - 1) * size);
}

/* Rust `f32 as u32`: saturating, NaN -> 0. */
static inline uint32_t f32_as_u32(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

/* Rust f32::clamp: NaN stays NaN. */
static inline float clampf_rust(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

/* fluid.rs:133-189 manage_single_cell_boundary */
static void manage_single_cell_boundary(int orientation, float *x, int64_t size, int64_t rows,
                                        const uint8_t *cells, int64_t i, int64_t j) {
    if (cells[IDX(i, j, size, rows)] == REF_DEFAULT_WALL) return;      /* :141-143 */

    const int64_t up_x = i, up_y = clampi(j - 1, 0, rows - 1);         /* :145 */
    const int64_t down_x = i, down_y = clampi(j + 1, 0, rows - 1);     /* :146 */
    const int64_t left_x = clampi(i - 1, 0, size - 1), left_y = j;     /* :147 */
    const int64_t right_x = clampi(i + 1, 0, size - 1), right_y = j;   /* :148 */

    switch (orientation) {
    case REF_ADJUST_ROW:                                                /* :151-164 */
        if (cells[IDX(left_x, left_y, size, rows)] == REF_DEFAULT_WALL)
            x[IDX(i, j, size, rows)] = -x[IDX(left_x, left_y, size, rows)];
        if (cells[IDX(right_x, right_y, size, rows)] == REF_DEFAULT_WALL)
            x[IDX(i, j, size, rows)] = -x[IDX(right_x, right_y, size, rows)];
        break;
    case REF_ADJUST_COLUMN:                                             /* :165-178 */
        if (cells[IDX(down_x, down_y, size, rows)] == REF_DEFAULT_WALL)
            x[IDX(i, j, size, rows)] = -x[IDX(down_x, down_y, size, rows)];
        if (cells[IDX(up_x, up_y, size, rows)] == REF_DEFAULT_WALL)
            x[IDX(i, j, size, rows)] = -x[IDX(up_x, up_y, size, rows)];
        break;
    default:                                                            /* :179-187 Passive */
        x[IDX(i, 0, size, rows)] = x[IDX(i, 1, size, rows)];
        x[IDX(i, rows - 1, size, rows)] = x[IDX(i, rows - 2, size, rows)];
        x[IDX(0, j, size, rows)] = x[IDX(1, j, size, rows)];
        x[IDX(size - 1, j, size, rows)] = x[IDX(size - 2, j, size, rows)];
        break;
    }
}

/* fluid.rs:252-272 set_boundaries */
void ref_set_boundaries(int orientation, float *x, uint32_t size_u, uint32_t rows_u,
                        const uint8_t *cells) {
    const int64_t size = (int64_t)size_u, rows = (int64_t)rows_u;
    for (int64_t j = 0; j <= rows - 1; ++j)                             /* :259 */
        for (int64_t i = 0; i <= size - 1; ++i)                         /* :260 */
            manage_single_cell_boundary(orientation, x, size, rows, cells, i, j);

    x[IDX(0, 0, size, rows)] =
        0.5f * (x[IDX(1, 0, size, rows)] + x[IDX(0, 1, size, rows)]);   /* :265 */
    x[IDX(0, rows - 1, size, rows)] =
        0.5f * (x[IDX(1, rows - 1, size, rows)] + x[IDX(0, rows - 2, size, rows)]); /* :266-267 */
    x[IDX(size - 1, 0, size, rows)] =
        0.5f * (x[IDX(size - 2, 0, size, rows)] + x[IDX(size - 1, 1, size, rows)]); /* :268-269 */
    x[IDX(size - 1, rows - 1, size, rows)] =
        0.5f * (x[IDX(size - 2, rows - 1, size, rows)] +
                x[IDX(size - 1, rows - 2, size, rows)]);                /* :270-271 */
}

/* fluid.rs:301-325 lin_solve: in-place lexicographic Gauss-Seidel */
void ref_lin_solve(int orientation, float *x, const float *x0, float a, float c,
                   uint32_t size, uint32_t rows, int64_t iters, const uint8_t *cells) {
    const float c_recip = 1.0f / c;                                     /* :311 */
    for (int64_t k = 0; k < iters; ++k) {                               /* :312 */
        for (uint32_t j = 1; j + 1 < rows; ++j) {                       /* :313 */
            for (uint32_t i = 1; i + 1 < size; ++i) {                   /* :314 */
                const size_t o = (size_t)i + (size_t)j * size;
                float s = x[o + 1] + x[o - 1];                          /* :316-317 */
                s = s + x[o + size];                                    /* :318 */
                s = s + x[o - size];                                    /* :319 */
                x[o] = (x0[o] + a * s) * c_recip;                       /* :315,:320 */
            }
        }
        ref_set_boundaries(orientation, x, size, rows, cells);          /* :323 */
    }
}

void ref_lin_solve_red_black(int orientation, float *x, const float *x0, float a, float c,
                             uint32_t size, uint32_t rows, int64_t iters,
                             const uint8_t *cells) {
    const float c_recip = 1.0f / c;
    for (int64_t k = 0; k < iters; ++k) {
        for (uint32_t colour = 0; colour < 2; ++colour) {
            for (uint32_t j = 1; j + 1 < rows; ++j) {
                for (uint32_t i = 1; i + 1 < size; ++i) {
                    if (((i + j) & 1u) != colour) continue;
                    const size_t o = (size_t)i + (size_t)j * size;
                    float s = x[o + 1] + x[o - 1];
                    s = s + x[o + size];
                    s = s + x[o - size];
                    x[o] = (x0[o] + a * s) * c_recip;
                }
            }
        }
        ref_set_boundaries(orientation, x, size, rows, cells);
    }
}

/* fluid.rs:276-298 diffuse */
void ref_diffuse(int orientation, float *x, const float *x0, float diffusion, uint32_t size,
                 uint32_t rows, float delta_t, int64_t iters, const uint8_t *cells) {
    const float size_float = (float)(size - 2);                         /* :286 */
    float a = delta_t * diffusion;                                      /* :287 (left to right) */
    a = a * size_float;
    a = a * size_float;
    ref_lin_solve(orientation, x, x0, a, 1.0f + 4.0f * a, size, rows, iters, cells); /* :288-297 */
}

/* fluid.rs:330-375 project */
void ref_project(float *vx, float *vy, float *p, float *div, uint32_t size, uint32_t rows,
                 int64_t iters, const uint8_t *cells) {
    const float nf = (float)size;
    for (uint32_t j = 1; j + 1 < rows; ++j) {                           /* :339 */
        for (uint32_t i = 1; i + 1 < size; ++i) {                       /* :340 */
            const size_t o = (size_t)i + (size_t)j * size;
            float t = vx[o + 1] - vx[o - 1];                            /* :342 */
            t = t + vy[o + size];                                       /* :343 */
            t = t - vy[o - size];                                       /* :344 */
            div[o] = (-0.5f * t) / nf;                                  /* :341,:345 */
            p[o] = 0.0f;                                                /* :347 */
        }
    }
    ref_set_boundaries(REF_PASSIVE, div, size, rows, cells);            /* :351 */
    ref_set_boundaries(REF_PASSIVE, p, size, rows, cells);              /* :352 */
    ref_lin_solve(REF_PASSIVE, p, div, 1.0f, 4.0f, size, rows, iters, cells); /* :353-362 */

    for (uint32_t j = 1; j + 1 < rows; ++j) {                           /* :364 */
        for (uint32_t i = 1; i + 1 < size; ++i) {                       /* :365 */
            const size_t o = (size_t)i + (size_t)j * size;
            vx[o] -= (0.5f * (p[o + 1] - p[o - 1])) * nf;               /* :366-367 */
            vy[o] -= (0.5f * (p[o + size] - p[o - size])) * nf;         /* :368-369 */
        }
    }
    ref_set_boundaries(REF_ADJUST_ROW, vx, size, rows, cells);          /* :373 */
    ref_set_boundaries(REF_ADJUST_COLUMN, vy, size, rows, cells);       /* :374 */
}

/* fluid.rs:378-432 advect (including the row-serial `break`, quirk Q4) */
void ref_advect(int orientation, float *d, const float *d0, const float *vx, const float *vy,
                uint32_t size, uint32_t rows, float delta_t, const uint8_t *cells) {
    const float delta_t_x = delta_t * (float)(size - 2);                /* :390 */
    const float delta_t_y = delta_t_x;                                  /* :391 */
    const float size_float = (float)size;                               /* :396 */
    const float rows_float = (float)rows;  /* == size_float in the reference */

    for (uint32_t j = 1; j + 1 < rows; ++j) {                           /* :398 */
        for (uint32_t i = 1; i + 1 < size; ++i) {                       /* :399 */
            const size_t o = (size_t)i + (size_t)j * size;
            float x = (float)i - delta_t_x * vx[o];                     /* :400 */
            float y = (float)j - delta_t_y * vy[o];                     /* :401 */
            x = clampf_rust(x, 0.5f, size_float - 1.0f);                /* :403 */
            y = clampf_rust(y, 0.5f, rows_float - 1.0f);                /* :404 */
            const float i0 = floorf(x), i1 = i0 + 1.0f;                 /* :406-407 */
            const float j0 = floorf(y), j1 = j0 + 1.0f;                 /* :409-410 */
            const float s1 = x - i0, s0 = 1.0f - s1;                    /* :412-413 */
            const float t1 = y - j0, t0 = 1.0f - t1;                    /* :414-415 */
            const uint32_t i0i = f32_as_u32(i0), i1i = f32_as_u32(i1);  /* :417 */
            const uint32_t j0i = f32_as_u32(j0), j1i = f32_as_u32(j1);  /* :418 */

            if (i1 >= size_float || j1 >= rows_float) {                 /* :420 */
                d[o] = d[o - 1];                                        /* :421 */
                break;                                                  /* :422 */
            }
            d[o] = s0 * (t0 * d0[IDX(i0i, j0i, size, rows)] + t1 * d0[IDX(i0i, j1i, size, rows)]) +
                   s1 * (t0 * d0[IDX(i1i, j0i, size, rows)] + t1 * d0[IDX(i1i, j1i, size, rows)]);
                                                                        /* :424-428 */
        }
    }
    ref_set_boundaries(orientation, d, size, rows, cells);              /* :431 */
}

/* fluid.rs:437-524 step */
void ref_fluid_step(ref_fluid *f) {
    const uint32_t n = f->size, r = f->rows;
    const int64_t k = f->gs_iterations ? f->gs_iterations : f->frames; /* :445 (Q1) */
    const uint8_t *c = f->cells_type;

    ref_diffuse(REF_ADJUST_ROW, f->velocities_x0, f->velocities_x, f->viscosity, n, r,
                f->delta_t, k, c);                                      /* :438-447 */
    ref_diffuse(REF_ADJUST_COLUMN, f->velocities_y0, f->velocities_y, f->viscosity, n, r,
                f->delta_t, k, c);                                      /* :448-457 */
    ref_project(f->velocities_x0, f->velocities_y0, f->velocities_x, f->velocities_y, n, r, k,
                c);                                                     /* :459-467 */
    ref_advect(REF_ADJUST_ROW, f->velocities_x, f->velocities_x0, f->velocities_x0,
               f->velocities_y0, n, r, f->delta_t, c);                  /* :469-478 */
    ref_advect(REF_ADJUST_COLUMN, f->velocities_y, f->velocities_y0, f->velocities_x0,
               f->velocities_y0, n, r, f->delta_t, c);                  /* :480-489 */
    ref_project(f->velocities_x, f->velocities_y, f->velocities_x0, f->velocities_y0, n, r, k,
                c);                                                     /* :491-499 */
    ref_diffuse(REF_PASSIVE, f->scratch_space, f->density, f->diffusion, n, r, f->delta_t, k,
                c);                                                     /* :501-510 */
    ref_advect(REF_PASSIVE, f->density, f->scratch_space, f->velocities_x, f->velocities_y, n,
               r, f->delta_t, c);                                       /* :512-521 */
    memcpy(f->scratch_space, f->density, sizeof(float) * (size_t)n * r); /* :523 */
}

/* fluid.rs:120-124 */
void ref_add_density(ref_fluid *f, uint32_t x, uint32_t y, float amount) {
    const size_t o = IDX(x, y, f->size, f->rows);
    f->density[o] += amount;
    f->scratch_space[o] += amount;
}

/* fluid.rs:127-131 */
void ref_add_velocity(ref_fluid *f, uint32_t x, uint32_t y, float ax, float ay) {
    const size_t o = IDX(x, y, f->size, f->rows);
    f->velocities_x[o] += ax;
    f->velocities_y[o] += ay;
}

/* fluid.rs:527-539 (requires size >= 20: `size / 2 - 10` underflows u32 below that) */
static void init_density(ref_fluid *f) {
    for (uint32_t j = 0; j < f->rows; ++j)
        for (uint32_t i = 0; i < f->size; ++i) ref_add_density(f, i, j, 0.0f);
    const uint32_t ci = f->size / 2, cj = f->rows / 2;
    for (uint32_t j = cj - 10; j <= cj + 10; ++j)
        for (uint32_t i = ci - 10; i <= ci + 10; ++i) ref_add_density(f, i, j, 0.9f);
}

/* fluid.rs:542-548 */
static void init_velocities(ref_fluid *f) {
    for (uint32_t j = 0; j < f->rows; ++j)
        for (uint32_t i = 0; i < f->size; ++i) ref_add_velocity(f, i, j, 1.0f, 1.0f);
}

/* fluid.rs:552-570 */
static void init_walls(ref_fluid *f) {
    for (uint32_t i = 0; i < f->size; ++i) {
        f->cells_type[IDX(i, 0, f->size, f->rows)] = REF_DEFAULT_WALL;
        f->cells_type[IDX(i, f->rows - 1, f->size, f->rows)] = REF_DEFAULT_WALL;
    }
    for (uint32_t j = 0; j < f->rows; ++j) {
        f->cells_type[IDX(0, j, f->size, f->rows)] = REF_DEFAULT_WALL;
        f->cells_type[IDX(f->size - 1, j, f->size, f->rows)] = REF_DEFAULT_WALL;
    }
}

/* fluid.rs:602-606 */
void ref_fluid_init(ref_fluid *f) {
    init_velocities(f);
    init_density(f);
    init_walls(f);
}

/* fluid.rs:93-110 */
ref_fluid *ref_fluid_new(uint32_t size, uint32_t rows, float delta_t, int64_t frames,
                         int64_t gs_iterations, float diffusion, float viscosity) {
    if (size < 20 || rows < 20) return NULL;
    ref_fluid *f = (ref_fluid *)calloc(1, sizeof(ref_fluid));
    if (!f) return NULL;
    const size_t n = (size_t)size * rows;
    f->size = size;
    f->rows = rows;
    f->delta_t = delta_t;
    f->frames = frames;
    f->gs_iterations = gs_iterations;
    f->diffusion = diffusion;
    f->viscosity = viscosity;
    f->scratch_space = (float *)calloc(n, sizeof(float));
    f->density = (float *)calloc(n, sizeof(float));
    f->velocities_x = (float *)calloc(n, sizeof(float));
    f->velocities_y = (float *)calloc(n, sizeof(float));
    f->velocities_x0 = (float *)calloc(n, sizeof(float));
    f->velocities_y0 = (float *)calloc(n, sizeof(float));
    f->cells_type = (uint8_t *)calloc(n, 1); /* NoWall */
    if (!f->scratch_space || !f->density || !f->velocities_x || !f->velocities_y ||
        !f->velocities_x0 || !f->velocities_y0 || !f->cells_type) {
        ref_fluid_free(f);
        return NULL;
    }
    ref_fluid_init(f);                                                  /* :108 */
    return f;
}

void ref_fluid_free(ref_fluid *f) {
    if (!f) return;
    free(f->scratch_space);
    free(f->density);
    free(f->velocities_x);
    free(f->velocities_y);
    free(f->velocities_x0);
    free(f->velocities_y0);
    free(f->cells_type);
    free(f);
}

/* fluid.rs:610-619: half-open ranges, idx! clamps out-of-range points */
void ref_fill_rect(ref_fluid *f, int64_t x0, int64_t y0, int64_t x1, int64_t y1) {
    for (int64_t x = x0; x < x1; ++x)
        for (int64_t y = y0; y < y1; ++y)
            f->cells_type[IDX(x, y, f->size, f->rows)] = REF_DEFAULT_WALL;
}

/* obstacle.rs:74-87 */
int ref_rect_valid(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t size) {
    return x0 != x1 && y0 != y1 && x0 < x1 && y0 < y1 && x0 < size && y0 < size && x1 < size &&
           y1 < size;
}

void *ref_fluid_field(ref_fluid *f, int id) {
    switch (id) {
    case REF_F_DENSITY: return f->density;
    case REF_F_VX: return f->velocities_x;
    case REF_F_VY: return f->velocities_y;
    case REF_F_VX0: return f->velocities_x0;
    case REF_F_VY0: return f->velocities_y0;
    case REF_F_SCRATCH: return f->scratch_space;
    case REF_F_CELLS: return f->cells_type;
    default: return NULL;
    }
}

/* Rust `f32 as u8`: truncates toward zero, saturates to 0..255, NaN -> 0. */
static inline uint8_t f32_as_u8(float v) {
    if (!(v > 0.0f)) return 0;          /* negatives, -0, NaN */
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

/* renderer_helpers.rs:145-167 */
void ref_render_rgba(const float *density, const uint8_t *cells, uint32_t size, uint32_t rows,
                     const uint8_t world[4], const uint8_t fluid[4], const uint8_t obstacle[4],
                     uint8_t *out) {
    for (uint32_t y = 0; y < rows; ++y)
        for (uint32_t x = 0; x < size; ++x) {
            const size_t o = (size_t)x + (size_t)y * size;
            const float d = density[o];
            uint8_t *px = out + 4 * o;
            if (cells[o]) {                                          /* :149-155 DefaultWall */
                px[0] = obstacle[0]; px[1] = obstacle[1]; px[2] = obstacle[2]; px[3] = obstacle[3];
            } else if (d != 0.0f) {                                  /* :156-162 */
                px[0] = f32_as_u8(d * (float)fluid[0]);
                px[1] = fluid[1];
                px[2] = f32_as_u8(d);
                px[3] = 1;
            } else {                                                 /* :163-165 */
                px[0] = world[0]; px[1] = world[1]; px[2] = world[2]; px[3] = world[3];
            }
        }
}

/* ---- device-side add_noise and dense sources ---------------------------
 * Philox4x32-10, restated from the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel
 * random numbers: as easy as 1, 2, 3", SC'11); pinned by the Random123 known-answer vectors in
 * tests/test_sources.py. */
void ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The structure of add_noise (fluid.rs:575-599) with a seeded draw: random grid point (:584-585),
 * rotated about the centre (:587-593; geo's rotate_around_point: x' = cos*(x-cx) - sin*(y-cy) + cx,
 * y' = sin*(x-cx) + cos*(y-cy) + cy), times gain (:595-596).  xy = centre cell, a = impulse. */
void ref_noise_impulse(uint64_t seed, uint64_t frame, uint32_t size, float cos_t, float sin_t, float gain,
                       uint32_t xy[2], float a[2]) {
    const uint32_t ctr[4] = {(uint32_t)frame, (uint32_t)(frame >> 32), 0u, 0u};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    ref_philox4x32_10(ctr, key, r);
    const uint32_t rx = (uint32_t)(((uint64_t)r[0] * size) >> 32), ry = (uint32_t)(((uint64_t)r[1] * size) >> 32);
    const float c = (float)(size / 2u);
    const float dx = (float)rx - c, dy = (float)ry - c;
    const float px = (cos_t * dx - sin_t * dy) + c;
    const float py = (sin_t * dx + cos_t * dy) + c;
    xy[0] = size / 2u; xy[1] = size / 2u;
    a[0] = px * gain; a[1] = py * gain;
}

/* Stam's add_source on a whole field: x += scale * s */
void ref_add_source(float *x, const float *s, float scale, size_t cells) {
    for (size_t i = 0; i < cells; ++i) x[i] = x[i] + scale * s[i];
}
