"""`CurrentSimulation`: the caller of the hot path, mirroring src/simulation/renderer_helpers.rs:29-81.

``simulate(tx)`` is the reference's loop -- mark the obstacles, then per frame ``add_noise`` (when enabled), ``step``
and a deep copy of the fluid sent to ``tx``.  ``simulate_frames(tx)`` is the same loop over the snapshot path
(SURVEY 8f rows 1-2): only the array the render thread consumes leaves the GPU, into pinned double buffers, while the
next step runs."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .fluid import Fluid
from .obstacle import Rectangle


@dataclass
class FluidStep:
    """renderer_helpers.rs:21-24"""
    fluid: Fluid
    frame_number: int


class CurrentSimulation:
    def __init__(self, fluid: Fluid | None = None, obstacles=None, **fluid_kw):
        """Default (renderer_helpers.rs:39-48): Fluid::default() and the default rectangle."""
        self.fluid = fluid if fluid is not None else Fluid.default(**fluid_kw)
        self.obstacles = list(obstacles) if obstacles is not None else [Rectangle.default()]

    def mark_fluid_obstacles(self):
        """renderer_helpers.rs:76-80"""
        for obstacle in self.obstacles:
            self.fluid.fill_obstacle(obstacle)

    def _advance(self):
        if self.fluid.fluid_configs.has_perlin_noise:
            self.fluid.add_noise()
        self.fluid.step()

    def simulate(self, tx):
        """renderer_helpers.rs:52-72; ``tx`` is any callable (the reference's mpsc Sender)."""
        self.mark_fluid_obstacles()
        for i in range(int(self.fluid.simulation_configs.frames)):
            self._advance()
            tx(FluidStep(self.fluid.clone(), i))

    def simulate_frames(self, tx, rgba: bool = False, obstacles_color=None):
        """``tx(frame_number, array)`` per frame, in order, each as soon as it has landed: the f32 density or, with
        ``rgba``, the finished pixels of render_image (renderer_helpers.rs:145-167).  The array is a view of a pinned
        buffer that is re-used two frames later."""
        from . import _lib
        f, lib = self.fluid, self.fluid._lib
        self.mark_fluid_obstacles()
        n = int(f.simulation_configs.size)
        r0, r1 = f.owned_rows()
        shape, dtype = ((r1 - r0, n, 4), np.uint8) if rgba else ((r1 - r0, n), np.float32)
        nbytes = (r1 - r0) * n * 4
        ptrs, bufs = [], []
        try:
            for _ in range(2):
                p = C.c_void_p()
                _lib.check(lib, lib.eq_host_alloc(C.byref(p), nbytes))
                ptrs.append(p)
                bufs.append(np.frombuffer((C.c_char * nbytes).from_address(p.value), dtype=dtype).reshape(shape))
            frames = int(f.simulation_configs.frames)
            for i in range(frames):
                self._advance()
                f.snapshot_begin(bufs[i & 1], slot=i & 1, rgba=rgba, obstacles_color=obstacles_color)
                if i > 0:
                    f.snapshot_wait(1 - (i & 1))
                    tx(i - 1, bufs[1 - (i & 1)])
            if frames > 0:
                f.snapshot_wait((frames - 1) & 1)
                tx(frames - 1, bufs[(frames - 1) & 1])
        finally:
            f.sync()
            bufs.clear()
            for p in ptrs:
                lib.eq_host_free(p)
