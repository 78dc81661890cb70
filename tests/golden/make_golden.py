"""Regenerates tests/golden/default_scene.json, default_scene_density_f16.npy, render_rgba.json and device_noise.json.

The reference is Rust and cannot be built or imported in the build image, and
its own tests pin no output of step(), so these vectors come from
the repo's CPU oracle (oracle/fluid_ref.c) AFTER the second, independently
written restatement (oracle/pyref.py) reproduced every hashed state bit for bit.
They pin the oracle against silent drift; they are not reference outputs.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import loader as O   # noqa: E402
from oracle import pyref as P    # noqa: E402
from parity import impulses      # noqa: E402

FIELDS = [("density", 0, "density"), ("velocities_x", 1, "velocities_x"), ("velocities_y", 2, "velocities_y"),
          ("velocities_x0", 3, "velocities_x0"), ("velocities_y0", 4, "velocities_y0"),
          ("scratch_space", 5, "scratch_space")]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(n, k, frames, rects, imp, checkpoints):
    c = O.RefFluid(n, 0.02, k)
    p = P.PyFluid(n, 0.02, k)
    for r in rects:
        c.fill_rect(*r)
        p.fill_rect(*r)
    out = {}
    for fr in range(frames):
        if imp:
            _, x, y, ax, ay = imp[fr]
            c.add_velocity(x, y, ax, ay)
            p.add_velocity(x, y, ax, ay)
        c.step()
        p.step()
        if fr + 1 in checkpoints:
            rec = {}
            for name, fid, pname in FIELDS:
                a = c.field(fid)
                b = getattr(p, pname).reshape(n, n)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, fr)
                rec[name] = sha(a)
            rec["density_sum"] = float(np.sum(c.density, dtype=np.float64))
            rec["vx_absmax"] = float(np.abs(c.vx).max())
            out[str(fr + 1)] = rec
    return c, out


WORLD, FLUID, OBSTACLE = (94, 146, 162, 128), (208, 88, 157, 220), (255, 0, 0, 255)   # configs.rs:56-57, Color32::RED


def render_numpy(density, cells, world=WORLD, fluid=FLUID, obstacle=OBSTACLE):
    """Second, vectorised restatement of render_image's pixel rule (renderer_helpers.rs:145-167) used to cross-check
    oracle/fluid_ref.c: ref_render_rgba before its hash is committed."""
    def as_u8(v):                                  # Rust `f32 as u8`: truncate, saturate, NaN -> 0
        v = np.nan_to_num(v.astype(np.float32), nan=0.0, posinf=255.0, neginf=0.0)
        return np.clip(np.trunc(v), 0, 255).astype(np.uint8)
    out = np.empty(density.shape + (4,), dtype=np.uint8)
    out[...] = np.array(world, dtype=np.uint8)
    nz = (density != 0) & (cells == 0)
    out[nz, 0] = as_u8(density * np.float32(fluid[0]))[nz]
    out[nz, 1] = fluid[1]
    out[nz, 2] = as_u8(density)[nz]
    out[nz, 3] = 1
    out[cells != 0] = np.array(obstacle, dtype=np.uint8)
    return out


def render_golden():
    """Pixels of the default scene after 16 frames (the committed density) with the default colours."""
    dens = np.load(os.path.join(HERE, "default_scene_density_f16.npy"))
    c = O.RefFluid(128, 0.02, 16)
    c.fill_rect(80, 80, 110, 110)
    a = O.render_rgba(dens, c.cells, WORLD, FLUID, OBSTACLE)
    b = render_numpy(dens, c.cells)
    assert np.array_equal(a, b)
    rec = {"note": "oracle-generated (C and numpy restatements of renderer_helpers.rs:145-167 agree); not reference output",
           "scene": "default_128_k16, frame 16", "world": WORLD, "fluid": FLUID, "obstacle": OBSTACLE,
           "sha256": sha(a), "obstacle_pixels": int((a == np.array(OBSTACLE, dtype=np.uint8)).all(axis=2).sum()),
           "fluid_pixels": int((a[..., 3] == 1).sum())}
    with open(os.path.join(HERE, "render_rgba.json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print("written", os.path.join(HERE, "render_rgba.json"))


def noise_golden():
    """Device-side add_noise: impulses and states of a seeded run, written only after the C and the
    numpy restatements of Philox4x32-10 + the impulse arithmetic + step() agree bit for bit."""
    import math
    n, k, frames, seed, first = 64, 4, 6, 0x5EED0FEED, 2**32 - 2      # the counter crosses 2^32 inside the run
    dt = 0.02
    angle = float(np.float32(math.sin(12.9898 * dt + 78.233 * dt) * 6.28 * 2.0))    # Fluid.noise_angle()
    cs, sn = float(np.float32(math.cos(math.radians(angle)))), float(np.float32(math.sin(math.radians(angle))))
    rects = [(10, 20, 30, 40)]
    c, p = O.RefFluid(n, dt, k), P.PyFluid(n, dt, k)
    for r in rects:
        c.fill_rect(*r)
        p.fill_rect(*r)
    imps, states = [], {}
    for fr in range(frames):
        a = O.noise_impulse(seed, first + fr, n, cs, sn, 2.0)
        b = P.noise_impulse(seed, first + fr, n, cs, sn, 2.0)
        assert a[:2] == b[:2] and np.float32(a[2]).tobytes() == np.float32(b[2]).tobytes() \
            and np.float32(a[3]).tobytes() == np.float32(b[3]).tobytes(), (fr, a, b)
        imps.append([a[0], a[1], float(a[2]).hex(), float(a[3]).hex()])
        c.add_velocity(*a)
        p.add_velocity(*a)
        c.step()
        p.step()
        if fr + 1 in (1, frames):
            rec = {}
            for name, fid, pname in FIELDS:
                x, y = c.field(fid), getattr(p, pname).reshape(n, n)
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (name, fr)
                rec[name] = sha(x)
            states[str(fr + 1)] = rec
    rec = {"note": "oracle-generated (C and numpy restatements agree); the reference's add_noise is unseeded and cannot "
                   "be reproduced, see SURVEY.md 8a row a12", "n": n, "k": k, "rects": rects, "seed": seed,
           "first_frame": first, "delta_t": dt, "angle_deg": float(angle).hex(), "cos_t": float(cs).hex(),
           "sin_t": float(sn).hex(), "gain": 2.0, "impulses": imps, "frames": states}
    with open(os.path.join(HERE, "device_noise.json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print("written", os.path.join(HERE, "device_noise.json"))


def main():
    render_only = "--render-only" in sys.argv
    if render_only:
        render_golden()
        return
    if "--noise-only" in sys.argv:
        noise_golden()
        return
    gold = {"note": "oracle-generated (both restatements agree); not reference outputs",
            "scenes": {}}
    rect = [(80, 80, 110, 110)]                     # obstacle.rs:47-51
    c, h = run(128, 16, 16, rect, None, {1, 4, 16})  # configs.rs:14-22 defaults, noise off
    gold["scenes"]["default_128_k16"] = {"n": 128, "k": 16, "rects": rect, "impulse_seed": None,
                                         "wall_cells": int(c.cells.sum()), "frames": h}
    np.save(os.path.join(HERE, "default_scene_density_f16.npy"), c.density.copy())
    c, h = run(128, 16, 8, rect, impulses(128, 8, seed=0), {2, 8})
    gold["scenes"]["default_128_k16_impulses"] = {"n": 128, "k": 16, "rects": rect, "impulse_seed": 0,
                                                  "wall_cells": int(c.cells.sum()), "frames": h}
    c, h = run(64, 5, 6, [(10, 20, 30, 40), (40, 5, 50, 60)], impulses(64, 6, seed=3), {3, 6})
    gold["scenes"]["small_64_k5_impulses"] = {"n": 64, "k": 5, "rects": [(10, 20, 30, 40), (40, 5, 50, 60)],
                                              "impulse_seed": 3, "wall_cells": int(c.cells.sum()), "frames": h}
    with open(os.path.join(HERE, "default_scene.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("written", os.path.join(HERE, "default_scene.json"))
    render_golden()
    noise_golden()


if __name__ == "__main__":
    main()
