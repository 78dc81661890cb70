"""oracle/pyref.py -- SECOND CPU RESTATEMENT (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

An independently written Python/numba transliteration of the reference solver
(`/root/reference/src/simulation/fluid.rs`).  Its only job is to cross-check
`oracle/fluid_ref.c`: the reference's own tests pin no value produced by
`step()` and no Rust toolchain exists here ("parity unpinned"), so the two
restatements must agree bit-for-bit (tests/test_oracle.py) before either is
trusted.  Nothing outside tests/ may import this module.

All arithmetic is float32 with explicit `np.float32` constants so that numba
(LLVM, no fast-math, no FMA contraction) evaluates the same expression trees as
rustc does for the reference.
"""
from __future__ import annotations

import numpy as np

try:  # numba makes 128^2 x 100 frames take seconds instead of minutes
    from numba import njit
except Exception:  # pragma: no cover - pure python fallback
    def njit(*a, **k):
        def deco(fn):
            return fn
        if a and callable(a[0]):
            return a[0]
        return deco

ROW, COL, PASSIVE = 0, 1, 2
F32 = np.float32


@njit(cache=True)
def _clampi(v, lo, hi):
    if v < lo:
        return lo
    if v > hi:
        return hi
    return v


@njit(cache=True)
def _idx(x, y, n):
    # fluid.rs:31-35
    return _clampi(x, 0, n - 1) + _clampi(y, 0, n - 1) * n


@njit(cache=True)
def set_boundaries(orientation, x, n, wall):
    """fluid.rs:252-272 + :133-189.  `x` is a flat float32 array, `wall` flat uint8."""
    for j in range(n):
        for i in range(n):
            if wall[_idx(i, j, n)] == 1:
                continue
            up = _idx(i, _clampi(j - 1, 0, n - 1), n)
            down = _idx(i, _clampi(j + 1, 0, n - 1), n)
            left = _idx(_clampi(i - 1, 0, n - 1), j, n)
            right = _idx(_clampi(i + 1, 0, n - 1), j, n)
            me = _idx(i, j, n)
            if orientation == 0:
                if wall[left] == 1:
                    x[me] = -x[left]
                if wall[right] == 1:
                    x[me] = -x[right]
            elif orientation == 1:
                if wall[down] == 1:
                    x[me] = -x[down]
                if wall[up] == 1:
                    x[me] = -x[up]
            else:
                x[_idx(i, 0, n)] = x[_idx(i, 1, n)]
                x[_idx(i, n - 1, n)] = x[_idx(i, n - 2, n)]
                x[_idx(0, j, n)] = x[_idx(1, j, n)]
                x[_idx(n - 1, j, n)] = x[_idx(n - 2, j, n)]
    half = np.float32(0.5)
    x[_idx(0, 0, n)] = half * (x[_idx(1, 0, n)] + x[_idx(0, 1, n)])
    x[_idx(0, n - 1, n)] = half * (x[_idx(1, n - 1, n)] + x[_idx(0, n - 2, n)])
    x[_idx(n - 1, 0, n)] = half * (x[_idx(n - 2, 0, n)] + x[_idx(n - 1, 1, n)])
    x[_idx(n - 1, n - 1, n)] = half * (x[_idx(n - 2, n - 1, n)] + x[_idx(n - 1, n - 2, n)])


@njit(cache=True)
def lin_solve(orientation, x, x0, a, c, n, iters, wall):
    """fluid.rs:301-325"""
    c_recip = np.float32(1.0) / c
    for _k in range(iters):
        for j in range(1, n - 1):
            for i in range(1, n - 1):
                s = x[(i + 1) + j * n] + x[(i - 1) + j * n]
                s = s + x[i + (j + 1) * n]
                s = s + x[i + (j - 1) * n]
                x[i + j * n] = (x0[i + j * n] + a * s) * c_recip
        set_boundaries(orientation, x, n, wall)


@njit(cache=True)
def diffuse(orientation, x, x0, diff, n, dt, iters, wall):
    """fluid.rs:276-298"""
    sf = np.float32(n - 2)
    a = dt * diff
    a = a * sf
    a = a * sf
    lin_solve(orientation, x, x0, a, np.float32(1.0) + np.float32(4.0) * a, n, iters, wall)


@njit(cache=True)
def project(vx, vy, p, div, n, iters, wall):
    """fluid.rs:330-375"""
    nf = np.float32(n)
    mhalf = np.float32(-0.5)
    half = np.float32(0.5)
    for j in range(1, n - 1):
        for i in range(1, n - 1):
            t = vx[(i + 1) + j * n] - vx[(i - 1) + j * n]
            t = t + vy[i + (j + 1) * n]
            t = t - vy[i + (j - 1) * n]
            div[i + j * n] = (mhalf * t) / nf
            p[i + j * n] = np.float32(0.0)
    set_boundaries(2, div, n, wall)
    set_boundaries(2, p, n, wall)
    lin_solve(2, p, div, np.float32(1.0), np.float32(4.0), n, iters, wall)
    for j in range(1, n - 1):
        for i in range(1, n - 1):
            vx[i + j * n] = vx[i + j * n] - (half * (p[(i + 1) + j * n] - p[(i - 1) + j * n])) * nf
            vy[i + j * n] = vy[i + j * n] - (half * (p[i + (j + 1) * n] - p[i + (j - 1) * n])) * nf
    set_boundaries(0, vx, n, wall)
    set_boundaries(1, vy, n, wall)


@njit(cache=True)
def _as_u32(v):
    # Rust `f32 as u32`: saturating, NaN -> 0
    if not (v > np.float32(0.0)):
        return 0
    if v >= np.float32(4294967296.0):
        return 4294967295
    return int(v)


@njit(cache=True)
def advect(orientation, d, d0, vx, vy, n, dt, wall):
    """fluid.rs:378-432 (with the row `break`)"""
    dtx = dt * np.float32(n - 2)
    dty = dtx
    nf = np.float32(n)
    one = np.float32(1.0)
    lo = np.float32(0.5)
    hi = nf - one
    for j in range(1, n - 1):
        for i in range(1, n - 1):
            x = np.float32(i) - dtx * vx[i + j * n]
            y = np.float32(j) - dty * vy[i + j * n]
            if x < lo:
                x = lo
            if x > hi:
                x = hi
            if y < lo:
                y = lo
            if y > hi:
                y = hi
            i0 = np.float32(np.floor(x))
            i1 = i0 + one
            j0 = np.float32(np.floor(y))
            j1 = j0 + one
            s1 = x - i0
            s0 = one - s1
            t1 = y - j0
            t0 = one - t1
            i0i = _as_u32(i0)
            i1i = _as_u32(i1)
            j0i = _as_u32(j0)
            j1i = _as_u32(j1)
            if i1 >= nf or j1 >= nf:
                d[i + j * n] = d[(i - 1) + j * n]
                break
            d[i + j * n] = s0 * (t0 * d0[_idx(i0i, j0i, n)] + t1 * d0[_idx(i0i, j1i, n)]) + s1 * (
                t0 * d0[_idx(i1i, j0i, n)] + t1 * d0[_idx(i1i, j1i, n)]
            )
    set_boundaries(orientation, d, n, wall)


class PyFluid:
    """fluid.rs:51-110 `Fluid` (numeric fields only)."""

    def __init__(self, size=128, delta_t=0.02, frames=16, diffusion=0.0, viscosity=0.001,
                 gs_iterations=0):
        n = int(size)
        self.size = n
        self.delta_t = F32(delta_t)
        self.frames = int(frames)
        self.gs_iterations = int(gs_iterations)
        self.diffusion = F32(diffusion)
        self.viscosity = F32(viscosity)
        z = lambda: np.zeros(n * n, dtype=np.float32)
        self.scratch_space, self.density = z(), z()
        self.velocities_x, self.velocities_y = z(), z()
        self.velocities_x0, self.velocities_y0 = z(), z()
        self.cells_type = np.zeros(n * n, dtype=np.uint8)
        self.init()

    # fluid.rs:602-606
    def init(self):
        n = self.size
        self.velocities_x += F32(1.0)              # :542-548
        self.velocities_y += F32(1.0)
        d = self.density.reshape(n, n)             # :527-539 (row = y)
        s = self.scratch_space.reshape(n, n)
        c = n // 2
        d[c - 10:c + 11, c - 10:c + 11] += F32(0.9)
        s[c - 10:c + 11, c - 10:c + 11] += F32(0.9)
        w = self.cells_type.reshape(n, n)          # :552-570
        w[0, :] = 1
        w[n - 1, :] = 1
        w[:, 0] = 1
        w[:, n - 1] = 1

    def add_velocity(self, x, y, ax, ay):          # :127-131
        o = _idx(int(x), int(y), self.size)
        self.velocities_x[o] += F32(ax)
        self.velocities_y[o] += F32(ay)

    def add_density(self, x, y, amount):           # :120-124
        o = _idx(int(x), int(y), self.size)
        self.density[o] += F32(amount)
        self.scratch_space[o] += F32(amount)

    def fill_rect(self, x0, y0, x1, y1):           # :610-619
        for x in range(x0, x1):
            for y in range(y0, y1):
                self.cells_type[_idx(x, y, self.size)] = 1

    def step(self):                                # :437-524
        n, dt, w = self.size, self.delta_t, self.cells_type
        k = self.gs_iterations if self.gs_iterations else self.frames
        diffuse(ROW, self.velocities_x0, self.velocities_x, self.viscosity, n, dt, k, w)
        diffuse(COL, self.velocities_y0, self.velocities_y, self.viscosity, n, dt, k, w)
        project(self.velocities_x0, self.velocities_y0, self.velocities_x, self.velocities_y,
                n, k, w)
        advect(ROW, self.velocities_x, self.velocities_x0, self.velocities_x0,
               self.velocities_y0, n, dt, w)
        advect(COL, self.velocities_y, self.velocities_y0, self.velocities_x0,
               self.velocities_y0, n, dt, w)
        project(self.velocities_x, self.velocities_y, self.velocities_x0, self.velocities_y0,
                n, k, w)
        diffuse(PASSIVE, self.scratch_space, self.density, self.diffusion, n, dt, k, w)
        advect(PASSIVE, self.density, self.scratch_space, self.velocities_x, self.velocities_y,
               n, dt, w)
        self.scratch_space = self.density.copy()


# ---- device-side add_noise: second, independent restatement ---------------------------------
def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11) on python ints."""
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def noise_impulse(seed, frame, n, cos_t, sin_t, gain=2.0):
    """Seeded add_noise (structure of fluid.rs:575-599) in numpy float32, every operation rounded separately."""
    f = np.float32
    r = philox4x32_10([frame & 0xFFFFFFFF, frame >> 32, 0, 0], [seed & 0xFFFFFFFF, seed >> 32])
    rx, ry = (r[0] * n) >> 32, (r[1] * n) >> 32
    c, cs, sn, g = f(n // 2), f(cos_t), f(sin_t), f(gain)
    dx, dy = f(rx) - c, f(ry) - c
    px = (cs * dx - sn * dy) + c
    py = (sn * dx + cs * dy) + c
    return n // 2, n // 2, float(px * g), float(py * g)
