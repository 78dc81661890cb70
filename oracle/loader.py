"""ctypes binding of oracle/libfluid_ref.so -- TEST INFRASTRUCTURE ONLY.

Importers allowed by the repo's rules: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfluid_ref.so")

ROW, COL, PASSIVE = 0, 1, 2
F_DENSITY, F_VX, F_VY, F_VX0, F_VY0, F_SCRATCH, F_CELLS = range(7)
FIELD_NAMES = ["density", "vx", "vy", "vx0", "vy0", "scratch", "cells"]

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_SO) or (
        os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "fluid_ref.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    fp = C.POINTER(C.c_float)
    u8p = C.POINTER(C.c_uint8)
    L.ref_fluid_new.restype = C.c_void_p
    L.ref_fluid_new.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_int64, C.c_int64,
                                C.c_float, C.c_float]
    L.ref_fluid_free.argtypes = [C.c_void_p]
    L.ref_fluid_init.argtypes = [C.c_void_p]
    L.ref_add_density.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float]
    L.ref_add_velocity.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    L.ref_fill_rect.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    L.ref_rect_valid.restype = C.c_int
    L.ref_rect_valid.argtypes = [C.c_int64] * 5
    L.ref_fluid_step.argtypes = [C.c_void_p]
    L.ref_fluid_field.restype = C.c_void_p
    L.ref_fluid_field.argtypes = [C.c_void_p, C.c_int]
    L.ref_set_boundaries.argtypes = [C.c_int, fp, C.c_uint32, C.c_uint32, u8p]
    L.ref_lin_solve.argtypes = [C.c_int, fp, fp, C.c_float, C.c_float, C.c_uint32, C.c_uint32,
                                C.c_int64, u8p]
    L.ref_lin_solve_red_black.argtypes = L.ref_lin_solve.argtypes
    L.ref_diffuse.argtypes = [C.c_int, fp, fp, C.c_float, C.c_uint32, C.c_uint32, C.c_float,
                              C.c_int64, u8p]
    L.ref_project.argtypes = [fp, fp, fp, fp, C.c_uint32, C.c_uint32, C.c_int64, u8p]
    L.ref_advect.argtypes = [C.c_int, fp, fp, fp, fp, C.c_uint32, C.c_uint32, C.c_float, u8p]
    L.ref_render_rgba.argtypes = [fp, u8p, C.c_uint32, C.c_uint32, u8p, u8p, u8p, u8p]
    u32p = C.POINTER(C.c_uint32)
    L.ref_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.ref_noise_impulse.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_float, C.c_float, C.c_float, u32p, fp]
    L.ref_add_source.argtypes = [fp, fp, C.c_float, C.c_size_t]
    _lib = L
    return L


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class RefFluid:
    """Oracle-side `Fluid` (fluid.rs:51-110). Fields are numpy views of the C arrays."""

    def __init__(self, size=128, delta_t=0.02, frames=16, diffusion=0.0, viscosity=0.001,
                 gs_iterations=0, rows=None):
        L = lib()
        self.size = int(size)
        self.rows = int(rows) if rows else int(size)
        self._h = L.ref_fluid_new(self.size, self.rows, delta_t, frames, gs_iterations,
                                  diffusion, viscosity)
        if not self._h:
            raise ValueError("ref_fluid_new failed (size < 20 or out of memory)")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ref_fluid_free(h)

    def field(self, fid) -> np.ndarray:
        ptr = lib().ref_fluid_field(self._h, fid)
        n = self.size * self.rows
        if fid == F_CELLS:
            buf = (C.c_uint8 * n).from_address(ptr)
            return np.frombuffer(buf, dtype=np.uint8).reshape(self.rows, self.size)
        buf = (C.c_float * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(self.rows, self.size)

    density = property(lambda s: s.field(F_DENSITY))
    vx = property(lambda s: s.field(F_VX))
    vy = property(lambda s: s.field(F_VY))
    vx0 = property(lambda s: s.field(F_VX0))
    vy0 = property(lambda s: s.field(F_VY0))
    scratch = property(lambda s: s.field(F_SCRATCH))
    cells = property(lambda s: s.field(F_CELLS))

    def init(self):
        lib().ref_fluid_init(self._h)

    def add_density(self, x, y, a):
        lib().ref_add_density(self._h, x, y, a)

    def add_velocity(self, x, y, ax, ay):
        lib().ref_add_velocity(self._h, x, y, ax, ay)

    def fill_rect(self, x0, y0, x1, y1):
        lib().ref_fill_rect(self._h, x0, y0, x1, y1)

    def step(self, n=1):
        for _ in range(n):
            lib().ref_fluid_step(self._h)


def set_boundaries(orientation, x, cells):
    rows, size = x.shape
    lib().ref_set_boundaries(orientation, _f(x), size, rows, _u(cells))


def lin_solve(orientation, x, x0, a, c, iters, cells, red_black=False):
    rows, size = x.shape
    fn = lib().ref_lin_solve_red_black if red_black else lib().ref_lin_solve
    fn(orientation, _f(x), _f(x0), a, c, size, rows, iters, _u(cells))


def diffuse(orientation, x, x0, diffusion, delta_t, iters, cells):
    rows, size = x.shape
    lib().ref_diffuse(orientation, _f(x), _f(x0), diffusion, size, rows, delta_t, iters, _u(cells))


def project(vx, vy, p, div, iters, cells):
    rows, size = vx.shape
    lib().ref_project(_f(vx), _f(vy), _f(p), _f(div), size, rows, iters, _u(cells))


def advect(orientation, d, d0, vx, vy, delta_t, cells):
    rows, size = d.shape
    lib().ref_advect(orientation, _f(d), _f(d0), _f(vx), _f(vy), size, rows, delta_t, _u(cells))


def rect_valid(x0, y0, x1, y1, size) -> bool:
    return bool(lib().ref_rect_valid(x0, y0, x1, y1, size))


def render_rgba(density, cells, world, fluid, obstacle) -> np.ndarray:
    """renderer_helpers.rs:145-167: (rows, size, 4) u8 pixels."""
    rows, size = density.shape
    out = np.empty((rows, size, 4), dtype=np.uint8)
    cols = [np.ascontiguousarray(np.array(c, dtype=np.uint8)) for c in (world, fluid, obstacle)]
    lib().ref_render_rgba(_f(np.ascontiguousarray(density)), _u(np.ascontiguousarray(cells)), size, rows,
                          _u(cols[0]), _u(cols[1]), _u(cols[2]), _u(out.reshape(-1)))
    return out


def philox4x32_10(ctr, key):
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    lib().ref_philox4x32_10(c, k, o)
    return list(o)


def noise_impulse(seed, frame, size, cos_t, sin_t, gain=2.0):
    """(x, y, ax, ay) of the seeded add_noise (structure of fluid.rs:575-599)."""
    xy, a = (C.c_uint32 * 2)(), (C.c_float * 2)()
    lib().ref_noise_impulse(seed, frame, size, cos_t, sin_t, gain, xy, a)
    return int(xy[0]), int(xy[1]), float(a[0]), float(a[1])


def add_source(x, s, scale):
    assert x.flags.c_contiguous and s.flags.c_contiguous and x.shape == s.shape
    lib().ref_add_source(_f(x), _f(s), scale, x.size)
