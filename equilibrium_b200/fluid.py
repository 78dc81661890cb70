"""`Fluid`: the host-side mirror of the reference's solver API
(src/simulation/fluid.rs:51-110, :437, :575, :610) over the CUDA C ABI.

Same names and argument meaning as the Rust type: ``Fluid.new(fluid_configs,
simulation_configs)``, ``Fluid.default()``, ``step()``, ``add_noise()``,
``fill_obstacle(obstacle)``, ``clone()``, public fields ``density``,
``velocities_x``, ``velocities_y``, ``cells_type``, ``fluid_configs``,
``simulation_configs``.  The state lives in HBM behind an opaque handle; the
public fields are numpy snapshots downloaded on access.
"""
from __future__ import annotations

import ctypes as C
import enum
import math

import numpy as np

from . import _lib
from ._lib import EqParams, EqProfile, EqSource, EquilibriumError
from .configs import FluidConfigs, SimulationConfigs


class ContainerWall(enum.IntEnum):
    """fluid.rs:11-17.  Values are this repo's u8 encoding, not Rust's discriminants."""

    NoWall = 0
    DefaultWall = 1


class Fluid:
    FIELDS = {"density": _lib.F_DENSITY, "velocities_x": _lib.F_VX, "velocities_y": _lib.F_VY,
              "velocities_x0": _lib.F_VX0, "velocities_y0": _lib.F_VY0,
              "scratch_space": _lib.F_SCRATCH, "cells_type": _lib.F_CELLS}

    def __init__(self, fluid_configs: FluidConfigs | None = None,
                 simulation_configs: SimulationConfigs | None = None, *,
                 mode: str = "exact", gs_iterations: int = 0, device: int = 0,
                 noise_seed: int = 0, lib_path: str | None = None, rank: int = 0, world: int = 1,
                 _handle=None):
        """Fluid::new (fluid.rs:93-110).  Extra keyword-only knobs the CUDA path adds:
        mode ('exact' | 'red_black'), gs_iterations (0 => `frames`, quirk Q1), device."""
        self._lib = _lib.load(lib_path)
        self.fluid_configs = (fluid_configs or FluidConfigs()).copy()
        self.simulation_configs = (simulation_configs or SimulationConfigs()).copy()
        self._mode = {"exact": _lib.MODE_EXACT, "red_black": _lib.MODE_RED_BLACK}[mode]
        self._gs_iterations = int(gs_iterations)
        self._device = int(device)
        self._rank, self._world = int(rank), int(world)
        self._rng = np.random.default_rng(noise_seed)
        self._pushed = None
        if _handle is not None:
            self._h = _handle
            self._pushed = self._params_tuple()
            return
        self._h = C.c_void_p()
        p = self._params()
        _lib.check(self._lib, self._lib.eq_create(C.byref(p), C.byref(self._h)))
        self._pushed = self._params_tuple()

    # -- construction helpers ------------------------------------------------
    @classmethod
    def new(cls, init_fluid: FluidConfigs, init_simulation: SimulationConfigs, **kw) -> "Fluid":
        return cls(init_fluid, init_simulation, **kw)

    @classmethod
    def default(cls, **kw) -> "Fluid":
        """Default::default (fluid.rs:83-89): new() and then init() a second time."""
        f = cls(FluidConfigs(), SimulationConfigs(), **kw)
        _lib.check(f._lib, f._lib.eq_init_default(f._h))
        return f

    def _params(self) -> EqParams:
        p = EqParams()
        p.size = int(self.simulation_configs.size)
        p.delta_t = float(self.simulation_configs.delta_t)
        p.frames = int(self.simulation_configs.frames)
        p.gs_iterations = self._gs_iterations
        p.diffusion = float(self.fluid_configs.diffusion)
        p.viscosity = float(self.fluid_configs.viscousity)
        p.mode = self._mode
        p.device = self._device
        p.rank, p.world = self._rank, self._world
        return p

    def _params_tuple(self):
        p = self._params()
        return (p.size, p.delta_t, p.frames, p.gs_iterations, p.diffusion, p.viscosity, p.mode)

    def _push_params(self):
        """The config structs are public and mutable in the reference; push edits lazily."""
        t = self._params_tuple()
        if t != self._pushed:
            p = self._params()
            _lib.check(self._lib, self._lib.eq_set_params(self._h, C.byref(p)))
            self._pushed = t

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.eq_destroy(h)

    __del__ = close

    # -- the reference's methods ----------------------------------------------
    def step(self):
        """Fluid::step (fluid.rs:437-524). Enqueued; reading a field synchronises."""
        self._push_params()
        _lib.check(self._lib, self._lib.eq_step(self._h))

    def step_n(self, n: int, sources=None):
        """n frames with optional point sources [(frame, x, y, dvx, dvy, ddensity), ...]
        -- the loop of CurrentSimulation::simulate (renderer_helpers.rs:54-66) without
        a host round trip per frame."""
        self._push_params()
        arr, cnt = None, 0
        if sources:
            srt = sorted(sources, key=lambda s: s[0])
            arr = (EqSource * len(srt))()
            for i, s in enumerate(srt):
                arr[i].frame, arr[i].x, arr[i].y = int(s[0]), int(s[1]), int(s[2])
                arr[i].d_vx, arr[i].d_vy = float(s[3]), float(s[4])
                arr[i].d_density = float(s[5]) if len(s) > 5 else 0.0
            cnt = len(srt)
        _lib.check(self._lib, self._lib.eq_step_n(self._h, int(n), arr, cnt))

    def add_velocity(self, x: int, y: int, amount_x: float, amount_y: float):
        """fluid.rs:127-131"""
        _lib.check(self._lib, self._lib.eq_add_velocity(self._h, x, y, amount_x, amount_y))

    def add_density(self, x: int, y: int, amount: float):
        """fluid.rs:120-124"""
        _lib.check(self._lib, self._lib.eq_add_density(self._h, x, y, amount))

    def noise_impulse(self):
        """The (x, y, ax, ay) that add_noise injects (fluid.rs:575-599).

        The reference draws the point from an unseeded thread_rng and the angle
        from noise-0.7 Perlin / geo-0.18 rotation; none of that is vendored or
        pinned by a test, so this is a seeded stand-in with the same structure:
        rotate a uniformly random grid point about the centre by a fixed angle
        (degrees, as geo's rotate_around_point takes) and add twice the rotated
        point as a velocity impulse at the centre cell."""
        n = int(self.simulation_configs.size)
        dt = float(self.simulation_configs.delta_t)
        angle = math.sin(12.9898 * dt + 78.233 * dt) * 6.28 * 2.0   # stands in for Perlin::get([dt,dt])
        rx, ry = int(self._rng.integers(0, n)), int(self._rng.integers(0, n))
        c = float(n // 2)
        th = math.radians(np.float32(angle))
        dx, dy = rx - c, ry - c
        px = c + dx * math.cos(th) - dy * math.sin(th)
        py = c + dx * math.sin(th) + dy * math.cos(th)
        return n // 2, n // 2, float(np.float32(px) * np.float32(2.0)), float(np.float32(py) * np.float32(2.0))

    def noise_angle(self) -> float:
        """The angle add_noise rotates by (fluid.rs:578-583): a function of delta_t only.  sin() stands in for
        noise-0.7's Perlin::get([dt, dt]) (not vendored, see noise_impulse)."""
        dt = float(self.simulation_configs.delta_t)
        return float(np.float32(math.sin(12.9898 * dt + 78.233 * dt) * 6.28 * 2.0))

    def device_noise(self, seed: int, first_frame: int = 0) -> "_lib.EqNoise":
        """Parameters of the device-side add_noise: Philox4x32-10 keyed by `seed`, one counter per
        frame; the rotation's cos/sin are evaluated once here (geo takes degrees), gain 2.0 (fluid.rs:595-596)."""
        th = math.radians(self.noise_angle())
        nz = _lib.EqNoise()
        nz.seed, nz.first_frame = int(seed) & (2**64 - 1), int(first_frame)
        nz.cos_t, nz.sin_t, nz.gain = math.cos(th), math.sin(th), 2.0
        return nz

    def step_n_noise(self, n: int, seed: int, first_frame: int = 0):
        """n x { add_noise(); step() } (renderer_helpers.rs:54-60) with the impulses drawn on the device."""
        self._push_params()
        nz = self.device_noise(seed, first_frame)
        _lib.check(self._lib, self._lib.eq_step_n_noise(self._h, int(n), C.byref(nz)))

    def add_noise_device(self, seed: int, frame: int):
        """One device-side add_noise (eq_add_noise): the impulse of Philox counter `frame`, no step."""
        nz = self.device_noise(seed, frame)
        _lib.check(self._lib, self._lib.eq_add_noise(self._h, C.byref(nz)))

    def add_noise(self):
        x, y, ax, ay = self.noise_impulse()
        self.add_velocity(x, y, ax, ay)

    def fill_obstacle(self, obstacle):
        """fluid.rs:610-619: mark [p0.x,p1.x) x [p0.y,p1.y) as DefaultWall."""
        (x0, y0), (x1, y1) = obstacle.get_approximate_points()[:2]
        _lib.check(self._lib, self._lib.eq_fill_rect(self._h, x0, y0, x1, y1))

    def reset_walls(self):
        _lib.check(self._lib, self._lib.eq_reset_walls(self._h))

    def clone(self) -> "Fluid":
        """#[derive(Clone)] (fluid.rs:51)"""
        self._push_params()
        h = C.c_void_p()
        _lib.check(self._lib, self._lib.eq_clone(self._h, C.byref(h)))
        f = Fluid(self.fluid_configs, self.simulation_configs, _handle=h,
                  mode="exact" if self._mode == _lib.MODE_EXACT else "red_black",
                  gs_iterations=self._gs_iterations, device=self._device)
        f._lib = self._lib
        return f

    def sync(self):
        _lib.check(self._lib, self._lib.eq_sync(self._h))

    # -- row slabs over several GPUs --------------------------------
    @property
    def rank(self) -> int:
        return self._rank

    @property
    def world(self) -> int:
        return self._world

    def ipc_blob(self) -> bytes:
        """This rank's rendezvous blob (eq_ipc_export)."""
        n = self._lib.eq_ipc_blob_bytes()
        buf = C.create_string_buffer(n)
        _lib.check(self._lib, self._lib.eq_ipc_export(self._h, buf, n))
        return buf.raw

    def ipc_attach(self, blobs):
        """Map every rank's arrays (eq_ipc_attach); `blobs` in rank order."""
        n = self._lib.eq_ipc_blob_bytes()
        assert len(blobs) == self._world and all(len(b) == n for b in blobs)
        joined = b"".join(blobs)
        _lib.check(self._lib, self._lib.eq_ipc_attach(self._h, joined, n, self._world))

    def owned_rows(self):
        a, b = C.c_uint32(), C.c_uint32()
        _lib.check(self._lib, self._lib.eq_owned_rows(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download_owned(self, name: str):
        """(row_begin, rows) of the slab this rank owns."""
        fid = self.FIELDS[name]
        n = int(self.simulation_configs.size)
        r0, r1 = self.owned_rows()
        dt = np.uint8 if fid == _lib.F_CELLS else np.float32
        out = np.empty((r1 - r0, n), dtype=dt)
        _lib.check(self._lib, self._lib.eq_download_rows(self._h, fid, r0, r1 - r0, out.ctypes.data))
        return r0, out

    # -- field access -----------------------------------------------------------
    def download(self, name: str, out: np.ndarray | None = None) -> np.ndarray:
        fid = self.FIELDS[name]
        n = int(self.simulation_configs.size)
        dt = np.uint8 if fid == _lib.F_CELLS else np.float32
        if out is None:
            out = np.empty((n, n), dtype=dt)
        assert out.dtype == dt and out.size == n * n and out.flags["C_CONTIGUOUS"]
        _lib.check(self._lib, self._lib.eq_download(self._h, fid, out.ctypes.data, out.nbytes))
        return out

    def upload(self, name: str, arr: np.ndarray):
        fid = self.FIELDS[name]
        dt = np.uint8 if fid == _lib.F_CELLS else np.float32
        a = np.ascontiguousarray(arr, dtype=dt)
        _lib.check(self._lib, self._lib.eq_upload(self._h, fid, a.ctypes.data, a.nbytes))

    # -- the caller's side of the frame loop (renderer_helpers.rs:61-65, 115-167) ----
    OBSTACLES_COLOR = (255, 0, 0, 255)     # RenderingListener::default, renderer_helpers.rs:94-101 (Color32::RED)

    def _colors(self, obstacles_color=None):
        c = _lib.EqColors()
        for dst, src in ((c.world, self.fluid_configs.world_color), (c.fluid, self.fluid_configs.fluid_color),
                         (c.obstacle, obstacles_color or self.OBSTACLES_COLOR)):
            for i in range(4):
                dst[i] = int(src[i])
        return c

    def render_rgba(self, obstacles_color=None, out: np.ndarray | None = None) -> np.ndarray:
        """render_image's pixel loop (renderer_helpers.rs:145-167) on the device: (owned rows, size, 4) u8."""
        n = int(self.simulation_configs.size)
        r0, r1 = self.owned_rows()
        if out is None:
            out = np.empty((r1 - r0, n, 4), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.nbytes == (r1 - r0) * n * 4 and out.flags["C_CONTIGUOUS"]
        c = self._colors(obstacles_color)
        _lib.check(self._lib, self._lib.eq_render_rgba(self._h, C.byref(c), out.ctypes.data, out.nbytes))
        return out

    def snapshot_begin(self, out: np.ndarray, slot: int = 0, rgba: bool = False, obstacles_color=None):
        """Asynchronous frame snapshot (what `fluid.clone()` + send is for the reference's render thread,
        renderer_helpers.rs:61-65): density (f32) or finished RGBA pixels into `out`, overlapping later steps.
        `out` should live in pinned memory (eq_host_alloc) and must stay alive until snapshot_wait(slot)."""
        c = self._colors(obstacles_color)
        kind = _lib.SNAP_RGBA if rgba else _lib.SNAP_DENSITY
        _lib.check(self._lib, self._lib.eq_snapshot_begin(self._h, kind, slot, C.byref(c), out.ctypes.data, out.nbytes))

    def snapshot_wait(self, slot: int = 0):
        _lib.check(self._lib, self._lib.eq_snapshot_wait(self._h, slot))

    density = property(lambda s: s.download("density"))
    velocities_x = property(lambda s: s.download("velocities_x"))
    velocities_y = property(lambda s: s.download("velocities_y"))
    cells_type = property(lambda s: s.download("cells_type"))

    # -- building blocks, profiling (used by tests and bench.py) ----------------
    def op_set_boundaries(self, orientation, field):
        _lib.check(self._lib, self._lib.eq_op_set_boundaries(self._h, orientation, self.FIELDS[field]))

    def op_lin_solve(self, orientation, x, x0, a, c, iters):
        self._push_params()
        _lib.check(self._lib, self._lib.eq_op_lin_solve(self._h, orientation, self.FIELDS[x], self.FIELDS[x0], a, c, iters))

    def op_diffuse(self, orientation, x, x0, diffusion, iters):
        self._push_params()
        _lib.check(self._lib, self._lib.eq_op_diffuse(self._h, orientation, self.FIELDS[x], self.FIELDS[x0], diffusion, iters))

    def op_project(self, vx, vy, p, div, iters):
        self._push_params()
        _lib.check(self._lib, self._lib.eq_op_project(self._h, self.FIELDS[vx], self.FIELDS[vy], self.FIELDS[p], self.FIELDS[div], iters))

    def op_add_source(self, x, s, scale):
        """x += scale * s over the whole grid (dense source field, Stam's add_source)."""
        _lib.check(self._lib, self._lib.eq_op_add_source(self._h, self.FIELDS[x], self.FIELDS[s], scale))

    def op_advect(self, orientation, d, d0, vx, vy):
        self._push_params()
        _lib.check(self._lib, self._lib.eq_op_advect(self._h, orientation, self.FIELDS[d], self.FIELDS[d0], self.FIELDS[vx], self.FIELDS[vy]))

    def divergence_l2(self, vx="velocities_x", vy="velocities_y") -> float:
        out = C.c_double()
        _lib.check(self._lib, self._lib.eq_divergence_l2(self._h, self.FIELDS[vx], self.FIELDS[vy], C.byref(out)))
        return out.value

    def timer_start(self):
        _lib.check(self._lib, self._lib.eq_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _lib.check(self._lib, self._lib.eq_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool):
        _lib.check(self._lib, self._lib.eq_profile_enable(self._h, int(on)))

    def profile_reset(self):
        _lib.check(self._lib, self._lib.eq_profile_reset(self._h))

    def profile(self) -> dict:
        p = EqProfile()
        _lib.check(self._lib, self._lib.eq_profile_get(self._h, C.byref(p)))
        return {k: getattr(p, k) for k, _ in EqProfile._fields_}

    def l2_flush(self):
        _lib.check(self._lib, self._lib.eq_l2_flush(self._h))

    def set_stream(self, cuda_stream_ptr: int | None):
        _lib.check(self._lib, self._lib.eq_set_stream(self._h, cuda_stream_ptr))


def connect_local(fluids):
    """Several row-slab handles living in ONE process (one per device): exchange blobs directly."""
    blobs = [f.ipc_blob() for f in fluids]
    for f in fluids:
        f.ipc_attach(blobs)


def connect_distributed(fluid: Fluid, group=None):
    """One process per GPU: gather the blobs with torch.distributed (plumbing only) and attach."""
    import torch.distributed as dist
    blobs = [None] * dist.get_world_size(group)
    dist.all_gather_object(blobs, fluid.ipc_blob(), group=group)
    fluid.ipc_attach(blobs)
    dist.barrier(group)
