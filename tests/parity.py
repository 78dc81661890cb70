"""Shared parity harness: drives the C ABI (real CUDA library or the emulated
build of the same sources) and the CPU oracle on identical seeded inputs and
compares bit patterns."""
from __future__ import annotations

import numpy as np

from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs

ROW, COL, PASSIVE = 0, 1, 2
ORIENT_NAMES = {ROW: "AdjustRow", COL: "AdjustColumn", PASSIVE: "Passive"}
F32_FIELDS = [("density", 0), ("velocities_x", 1), ("velocities_y", 2),
              ("velocities_x0", 3), ("velocities_y0", 4), ("scratch_space", 5)]


def random_rects(n: int, count: int, seed: int):
    """x0,y0 ~ U[1, N-2-w], w,h ~ U[N/64, N/16] (at least 1), all valid."""
    rng = np.random.default_rng(seed)
    out = []
    lo, hi = max(1, n // 64), max(2, n // 16)
    for _ in range(count):
        w, h = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        x0, y0 = int(rng.integers(1, n - 2 - w + 1)), int(rng.integers(1, n - 2 - h + 1))
        out.append((x0, y0, x0 + w, y0 + h))
    return out


def _canon(a: np.ndarray) -> np.ndarray:
    """uint32 view with every NaN mapped to one pattern: NaN *payload and sign* are the only
    bits not compared (x86 SSE produces 0xFFC00000 for an invalid operation, the GPU
    0x7FFFFFFF; IEEE 754 leaves the payload open and the reference never inspects it)."""
    u = np.ascontiguousarray(a).view(np.uint32).copy()
    u[np.isnan(a)] = 0x7FC00000
    return u


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return np.array_equal(_canon(a), _canon(b))


def describe_diff(a, b):
    bad = np.argwhere(_canon(a) != _canon(b))
    return f"{len(bad)} cells differ, first (row, col): {bad[:5].tolist()}"


def make_pair(oracle, lib_path, n, k, rects=(), *, frames=None, diffusion=0.0, viscosity=0.001,
              dt=0.02, mode="exact"):
    """A device fluid and an oracle fluid in the same state (Fluid::new + rectangles)."""
    frames = k if frames is None else frames
    gs = 0 if frames == k else k
    dev = Fluid(FluidConfigs(diffusion=diffusion, viscousity=viscosity),
                SimulationConfigs(dt, frames, n), lib_path=lib_path, mode=mode, gs_iterations=gs)
    ref = oracle.RefFluid(n, dt, frames, diffusion, viscosity, gs_iterations=gs)
    for (x0, y0, x1, y1) in rects:
        dev.fill_obstacle(Rectangle((x0, y0), (x1, y1), n))
        ref.fill_rect(x0, y0, x1, y1)
    return dev, ref


def assert_state_equal(dev, ref, where=""):
    assert np.array_equal(dev.download("cells_type"), ref.cells), f"{where}: cells_type differs"
    for name, fid in F32_FIELDS:
        a, b = dev.download(name), ref.field(fid)
        assert bits_equal(a, b), f"{where}: {name}: {describe_diff(a, b)}"


def rnd(rng, n, scale=1.0):
    return (rng.standard_normal((n, n)) * scale).astype(np.float32)


def check_set_boundaries(oracle, lib_path, n, rects, orient, seed=0):
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, 1, rects)
    x = rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.op_set_boundaries(orient, "velocities_x")
    oracle.set_boundaries(orient, x, ref.cells)
    got = dev.download("velocities_x")
    assert bits_equal(got, x), f"set_boundaries {ORIENT_NAMES[orient]} N={n}: {describe_diff(got, x)}"


def check_lin_solve(oracle, lib_path, n, k, rects, orient, a=0.37, seed=0):
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, k, rects)
    x, x0 = rnd(rng, n), rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    c = float(np.float32(1.0) + np.float32(4.0) * np.float32(a))
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", a, c, k)
    oracle.lin_solve(orient, x, x0, a, c, k, ref.cells)
    got = dev.download("velocities_x")
    assert bits_equal(got, x), f"lin_solve {ORIENT_NAMES[orient]} N={n} K={k}: {describe_diff(got, x)}"


def check_lin_solve_a0(oracle, lib_path, n, k, rects, orient, poison, seed=0, mode="exact"):
    """lin_solve with a == 0, c == 1 (diffuse with a zero coefficient, fluid.rs:286-297): the library answers with one
    guarded copy + set_boundaries when every value is finite, moderate and no x0 is -0.0, and falls back to the sweeps
    otherwise.  `poison` = None (the shortcut applies) or (array, row, col, value): one value that must defeat the guard."""
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, k, rects, mode=mode)
    x, x0 = rnd(rng, n), rnd(rng, n)
    x0[rng.random((n, n)) < 0.1] = 0.0                      # +0.0 is harmless
    if poison is not None:
        which, j, i, v = poison
        (x if which == "x" else x0)[j, i] = np.float32(v)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.0, 1.0, k)
    if mode == "exact":
        oracle.lin_solve(orient, x, x0, 0.0, 1.0, k, ref.cells)
    else:
        oracle.lin_solve(orient, x, x0, 0.0, 1.0, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    nan_both = np.isnan(got) & np.isnan(x)                  # NaN payloads are not compared (x86 vs GPU quiet-NaN patterns)
    assert np.array_equal(np.isnan(got), np.isnan(x)), "NaN cells differ"
    g, w = got.copy(), x.copy()
    g[nan_both] = 0.0
    w[nan_both] = 0.0
    assert bits_equal(g, w), f"lin_solve a=0 {ORIENT_NAMES[orient]} N={n} K={k} poison={poison}: {describe_diff(g, w)}"


def check_project(oracle, lib_path, n, k, rects, seed=0):
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, k, rects)
    names = ["velocities_x", "velocities_y", "velocities_x0", "velocities_y0"]
    arrs = [rnd(rng, n) for _ in names]
    for nm, a in zip(names, arrs):
        dev.upload(nm, a)
    dev.op_project(*names, k)
    oracle.project(*arrs, k, ref.cells)
    for nm, a in zip(names, arrs):
        got = dev.download(nm)
        assert bits_equal(got, a), f"project N={n} K={k} {nm}: {describe_diff(got, a)}"


def check_advect(oracle, lib_path, n, rects, orient, vscale=3.0, seed=0):
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, 1, rects)
    d, d0 = rnd(rng, n), rnd(rng, n)
    vx, vy = rnd(rng, n, vscale), rnd(rng, n, vscale)
    for nm, a in zip(["density", "scratch_space", "velocities_x", "velocities_y"], [d, d0, vx, vy]):
        dev.upload(nm, a)
    dev.op_advect(orient, "density", "scratch_space", "velocities_x", "velocities_y")
    oracle.advect(orient, d, d0, vx, vy, 0.02, ref.cells)
    got = dev.download("density")
    assert bits_equal(got, d), f"advect {ORIENT_NAMES[orient]} N={n}: {describe_diff(got, d)}"


def impulses(n, frames, seed=0):
    """scripted stand-in for add_noise: (N/2, N/2, U(-2N,2N), U(-2N,2N)) per frame."""
    rng = np.random.default_rng(seed)
    return [(fr, n // 2, n // 2, float(np.float32(rng.uniform(-2 * n, 2 * n))),
             float(np.float32(rng.uniform(-2 * n, 2 * n)))) for fr in range(frames)]


def check_steps(oracle, lib_path, n, k, frames_to_run, rects, *, with_impulses, diffusion=0.0,
                every_frame=True, use_step_n=False, seed=0):
    dev, ref = make_pair(oracle, lib_path, n, k, rects, diffusion=diffusion)
    imp = impulses(n, frames_to_run, seed) if with_impulses else []
    if use_step_n:
        dev.step_n(frames_to_run, imp)
        for (fr, x, y, ax, ay) in imp or [(f, 0, 0, 0, 0) for f in range(frames_to_run)]:
            if imp:
                ref.add_velocity(x, y, ax, ay)
            ref.step()
        assert_state_equal(dev, ref, f"N={n} K={k} after step_n({frames_to_run})")
        return dev, ref
    for fr in range(frames_to_run):
        if imp:
            _, x, y, ax, ay = imp[fr]
            dev.add_velocity(x, y, ax, ay)
            ref.add_velocity(x, y, ax, ay)
        dev.step()
        ref.step()
        if every_frame or fr == frames_to_run - 1:
            assert_state_equal(dev, ref, f"N={n} K={k} frame {fr}")
    return dev, ref


def check_render_and_snapshot(oracle, lib_path, n, rects, seed=0, steps=2):
    """The caller's side of the frame loop: the device colour map against the
    restatement of render_image's pixel loop (renderer_helpers.rs:145-167) on adversarial densities,
    then double-buffered snapshots taken between steps against plain downloads."""
    rng = np.random.default_rng(seed)
    dev, ref = make_pair(oracle, lib_path, n, 1, rects)
    d = (rng.standard_normal((n, n)) * 0.8).astype(np.float32)
    flat = d.reshape(-1)
    # zero / negative zero (world colour), saturation of `as u8` both ways, NaN (-> 0), infinities, denormals
    special = np.array([0.0, -0.0, 1.0, 254.99, 255.0, 256.0, 1e9, -3.0, np.nan, np.inf, -np.inf, 1e-40,
                        1.2259, 0.9999999, 1.0000001], dtype=np.float32)
    pos = rng.choice(flat.size, size=special.size * 8, replace=False)
    flat[pos] = np.tile(special, 8)
    dev.upload("density", d)
    world, fluid, obstacle = dev.fluid_configs.world_color, dev.fluid_configs.fluid_color, (255, 0, 0, 255)
    want = oracle.render_rgba(d, ref.cells, world, fluid, obstacle)
    got = dev.render_rgba(obstacle)
    assert got.shape == (n, n, 4)
    assert np.array_equal(got, want), f"render_rgba N={n}: {int((got != want).any(axis=2).sum())} pixels differ"
    if steps <= 0:
        return
    # snapshots: begin after a step, wait one step later (the copy is in flight while the state moves on); the
    # expectation is a plain download taken at the point of the snapshot
    bufs = [np.empty((n, n), dtype=np.float32), np.empty((n, n, 4), dtype=np.uint8)]
    dev.upload("density", np.nan_to_num(d, nan=0.5, posinf=2.0, neginf=-2.0))
    for s in range(steps):
        dev.step()
        rgba = bool(s & 1)
        dev.snapshot_begin(bufs[s & 1], slot=s & 1, rgba=rgba, obstacles_color=obstacle)
        dens = dev.download("density")
        dev.step()                                   # the state moves on while the snapshot is in flight
        dev.snapshot_wait(s & 1)
        if rgba:
            assert np.array_equal(bufs[1], oracle.render_rgba(dens, ref.cells, world, fluid, obstacle)), f"RGBA snapshot {s}"
        else:
            assert bits_equal(bufs[0], dens), f"density snapshot {s}: {describe_diff(bufs[0], dens)}"
        assert not bits_equal(dens, dev.download("density")), "the second step did not change the state"


def golden_render_case():
    """(density, rects, colours, record) of tests/golden/render_rgba.json: the default scene after 16 frames."""
    import json
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(gold, "render_rgba.json")) as f:
        rec = json.load(f)
    dens = np.load(os.path.join(gold, "default_scene_density_f16.npy"))
    return dens, [(80, 80, 110, 110)], rec


def check_golden_render(lib_path):
    """Device colour map of the committed default-scene density against the committed pixel hash (no oracle needed)."""
    import hashlib
    dens, rects, rec = golden_render_case()
    n = dens.shape[0]
    dev = Fluid(FluidConfigs(), SimulationConfigs(0.02, 16, n), lib_path=lib_path)
    for (x0, y0, x1, y1) in rects:
        dev.fill_obstacle(Rectangle((x0, y0), (x1, y1), n))
    dev.upload("density", dens)
    px = dev.render_rgba(tuple(rec["obstacle"]))
    assert hashlib.sha256(np.ascontiguousarray(px).tobytes()).hexdigest() == rec["sha256"]
    assert int((px == np.array(rec["obstacle"], dtype=np.uint8)).all(axis=2).sum()) == rec["obstacle_pixels"] == 1408


def check_device_noise(oracle, lib_path, n, k, frames_to_run, rects, seed, first_frame=0):
    """n x { device-side add_noise; step } against the oracle fed its own restatement of the
    impulse through add_velocity (fluid.rs:575-599 -> :127-131)."""
    dev, ref = make_pair(oracle, lib_path, n, k, rects)
    nz = dev.device_noise(seed, first_frame)
    dev.step_n_noise(frames_to_run, seed, first_frame)
    for fr in range(frames_to_run):
        x, y, ax, ay = oracle.noise_impulse(seed, first_frame + fr, n, nz.cos_t, nz.sin_t, nz.gain)
        assert (x, y) == (n // 2, n // 2) and abs(ax) <= 4 * n and abs(ay) <= 4 * n
        ref.add_velocity(x, y, ax, ay)
        ref.step()
    dev.sync()
    assert_state_equal(dev, ref, f"device noise N={n} K={k} frames={frames_to_run} seed={seed}")
    dev.close()


def check_add_source(oracle, lib_path, n, scale=0.02, seed=0):
    rng = np.random.default_rng(seed)
    dev, _ = make_pair(oracle, lib_path, n, 1)
    x, s = rnd(rng, n), rnd(rng, n, 5.0)
    other = rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", s)
    dev.upload("velocities_y", other)
    dev.op_add_source("velocities_x", "velocities_x0", scale)
    oracle.add_source(x, s, scale)
    got = dev.download("velocities_x")
    assert bits_equal(got, x), f"add_source N={n}: {describe_diff(got, x)}"
    assert bits_equal(dev.download("velocities_x0"), s) and bits_equal(dev.download("velocities_y"), other)
    dev.close()


def check_golden_device_noise(lib_path):
    """tests/golden/device_noise.json with no oracle in the loop: a seeded run whose Philox counter crosses 2^32."""
    import hashlib
    import json
    import os
    import ctypes as C
    from equilibrium_b200 import _lib
    with open(os.path.join(os.path.dirname(__file__), "golden", "device_noise.json")) as f:
        g = json.load(f)
    n, k = g["n"], g["k"]
    dev = Fluid(FluidConfigs(), SimulationConfigs(g["delta_t"], k, n), lib_path=lib_path)
    for r in g["rects"]:
        dev.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
    # the mirror derives the same rotation from delta_t as the generator did
    nz = dev.device_noise(g["seed"], g["first_frame"])
    assert float(nz.cos_t).hex() == g["cos_t"] and float(nz.sin_t).hex() == g["sin_t"] and nz.gain == g["gain"]
    done = 0
    for upto in sorted(int(s) for s in g["frames"]):
        nz.first_frame = g["first_frame"] + done
        _lib.check(dev._lib, dev._lib.eq_step_n_noise(dev._h, upto - done, C.byref(nz)))
        done = upto
        for name, _ in F32_FIELDS:
            got = hashlib.sha256(np.ascontiguousarray(dev.download(name)).tobytes()).hexdigest()
            assert got == g["frames"][str(upto)][name], (upto, name)
    dev.close()


def check_current_simulation(oracle, lib_path, n=64, k=3, rect=(20, 24, 40, 44), seed=5):
    """CurrentSimulation (renderer_helpers.rs:29-81): the reference's clone-per-frame loop, then the same run over the
    snapshot path (f32 density, then RGBA), each frame compared with the oracle."""
    from equilibrium_b200 import CurrentSimulation

    def make():
        f = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=lib_path, noise_seed=seed)
        twin = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=lib_path, noise_seed=seed)   # same host RNG
        ref = oracle.RefFluid(n, 0.02, k)
        ref.fill_rect(*rect)
        return CurrentSimulation(f, [Rectangle(rect[:2], rect[2:], n)]), twin, ref

    def ref_frame(twin, ref):
        ref.add_velocity(*twin.noise_impulse())
        ref.step()

    sim, twin, ref = make()
    seen = []

    def on_step(step):
        assert step.frame_number == len(seen)
        ref_frame(twin, ref)
        assert_state_equal(step.fluid, ref, f"FluidStep {step.frame_number}")
        seen.append(step.frame_number)
        step.fluid.close()

    sim.simulate(on_step)
    assert seen == list(range(k))

    for rgba in (False, True):
        sim, twin, ref = make()
        seen = []
        world, fluid = sim.fluid.fluid_configs.world_color, sim.fluid.fluid_configs.fluid_color

        def on_frame(i, arr):
            assert i == len(seen)
            ref_frame(twin, ref)
            if rgba:
                assert np.array_equal(arr, oracle.render_rgba(ref.density, ref.cells, world, fluid, (255, 0, 0, 255))), i
            else:
                assert bits_equal(arr, ref.density), f"frame {i}: {describe_diff(arr, ref.density)}"
            seen.append(i)

        sim.simulate_frames(on_frame, rgba=rgba)
        assert seen == list(range(k))
    # Default: Fluid::default() + the default rectangle, and has_perlin_noise = false skips add_noise
    sim = CurrentSimulation(lib_path=lib_path)
    assert len(sim.obstacles) == 1 and sim.fluid.simulation_configs.frames == 16
