"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs -- bit-exact in exact mode.

Sizes: the oracle finishes each case in seconds.  At the BASELINE sizes (4096^2,
16384^2) single operators are compared directly with reduced iteration counts and
the full-K runs are covered by size-independent properties (exact power-of-two
linearity, schedule independence)."""
import hashlib
import json
import os

import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n", [128, 200])
def test_set_boundaries(oracle, cuda_lib, orient, n):
    P.check_set_boundaries(oracle, cuda_lib, n, P.random_rects(n, 6, n), orient)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,k,nrect", [(64, 3, 3), (128, 16, 4), (333, 7, 9), (1024, 20, 0), (1024, 4, 32)])
def test_lin_solve_exact(oracle, cuda_lib, orient, n, k, nrect):
    P.check_lin_solve(oracle, cuda_lib, n, k, P.random_rects(n, nrect, n + k), orient)


def test_lin_solve_full_row_and_column_walls(oracle, cuda_lib):
    # quirk Q6: an obstacle spanning a whole interior row / column switches off the
    # Passive frame copy for it; also walls two cells apart (left and right both walls)
    n = 96
    rects = [(1, 40, 95, 41), (50, 1, 51, 95), (10, 10, 11, 30), (12, 10, 13, 30)]
    for orient in (P.ROW, P.COL, P.PASSIVE):
        P.check_lin_solve(oracle, cuda_lib, n, 5, rects, orient)
        P.check_set_boundaries(oracle, cuda_lib, n, rects, orient)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("poison", [None, ("x0", 200, 300, -0.0), ("x0", 5, 5, float("nan")), ("x", 0, 7, float("inf")),
                                    ("x", 400, 41, -3e38), ("x0", 1, 1, 2e37)])
def test_lin_solve_zero_coefficient_shortcut_and_its_guard(oracle, cuda_lib, orient, poison):
    # a = 0, c = 1 (the reference's default diffusion, configs.rs:50-60): guarded copy, or the sweeps when the guard fails
    P.check_lin_solve_a0(oracle, cuda_lib, 512, 6, P.random_rects(512, 8, 3), orient, poison)


def test_lin_solve_zero_coefficient_red_black(oracle, cuda_lib):
    P.check_lin_solve_a0(oracle, cuda_lib, 700, 9, P.random_rects(700, 8, 5), P.PASSIVE, None, mode="red_black")
    P.check_lin_solve_a0(oracle, cuda_lib, 700, 9, P.random_rects(700, 8, 5), P.PASSIVE, ("x0", 333, 21, -0.0), mode="red_black")


def test_lin_solve_more_iterations_than_one_launch(oracle, cuda_lib):
    # > LSX_KMAX (256) iterations are split over several wavefront launches
    P.check_lin_solve(oracle, cuda_lib, 64, 300, [(10, 10, 30, 20)], P.COL)


def test_lin_solve_zero_iterations(oracle, cuda_lib):
    P.check_lin_solve(oracle, cuda_lib, 64, 0, [], P.PASSIVE)


@pytest.mark.parametrize("n,k,nrect", [(128, 16, 4), (500, 5, 12)])
def test_project(oracle, cuda_lib, n, k, nrect):
    P.check_project(oracle, cuda_lib, n, k, P.random_rects(n, nrect, 7))


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,vscale", [(128, 4.0), (777, 0.5), (1024, 30.0)])
def test_advect(oracle, cuda_lib, orient, n, vscale):
    P.check_advect(oracle, cuda_lib, n, P.random_rects(n, 5, 11), orient, vscale=vscale)


def test_advect_nan_and_huge_velocities(oracle, cuda_lib):
    # `as u32` saturation / NaN -> 0 and f32::clamp keeping NaN (fluid.rs:403-418)
    rng = np.random.default_rng(5)
    n = 64
    dev, ref = P.make_pair(oracle, cuda_lib, n, 1, [(20, 20, 30, 30)])
    d, d0 = P.rnd(rng, n), P.rnd(rng, n)
    vx, vy = P.rnd(rng, n, 2.0), P.rnd(rng, n, 2.0)
    vx[10, 10], vy[11, 11], vx[12, 40], vy[13, 41] = np.nan, np.nan, 1e30, -1e30
    vx[30, 5], vy[31, 6] = np.inf, -np.inf
    for nm, a in zip(["density", "scratch_space", "velocities_x", "velocities_y"], [d, d0, vx, vy]):
        dev.upload(nm, a)
    dev.op_advect(P.PASSIVE, "density", "scratch_space", "velocities_x", "velocities_y")
    oracle.advect(P.PASSIVE, d, d0, vx, vy, 0.02, ref.cells)
    got = dev.download("density")
    assert P.bits_equal(got, d), P.describe_diff(got, d)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("scene", ["default_128_k16", "default_128_k16_impulses", "small_64_k5_impulses"])
def test_golden_scenes(cuda_lib, scene):
    """BASELINE config 1 (default scene) against the committed golden hashes -- no oracle
    in the loop at all."""
    from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs
    with open(os.path.join(GOLD, "default_scene.json")) as f:
        g = json.load(f)["scenes"][scene]
    n, k = g["n"], g["k"]
    dev = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=cuda_lib)
    for r in g["rects"]:
        dev.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
    assert int(dev.cells_type.sum()) == g["wall_cells"]
    last = max(int(s) for s in g["frames"])
    imp = P.impulses(n, last, g["impulse_seed"]) if g["impulse_seed"] is not None else None
    for fr in range(last):
        if imp:
            _, x, y, ax, ay = imp[fr]
            dev.add_velocity(x, y, ax, ay)
        dev.step()
        rec = g["frames"].get(str(fr + 1))
        if rec:
            for name in ["density", "velocities_x", "velocities_y", "velocities_x0", "velocities_y0",
                         "scratch_space"]:
                assert _sha(dev.download(name)) == rec[name], (scene, fr + 1, name)


def test_config1_default_scene_100_frames(oracle, cuda_lib):
    # BASELINE config 1: default grid, one rectangle, frames = 100 => 100 GS iterations (Q1)
    P.check_steps(oracle, cuda_lib, 128, 100, 12, [(80, 80, 110, 110)], with_impulses=False, every_frame=False)
    P.check_steps(oracle, cuda_lib, 128, 16, 16, [(80, 80, 110, 110)], with_impulses=True, every_frame=True)


def test_config2_1024_k20(oracle, cuda_lib):
    # BASELINE config 2: 1024^2, no obstacles, 20 GS iterations
    P.check_steps(oracle, cuda_lib, 1024, 20, 2, [], with_impulses=True, every_frame=True)


def test_ragged_size_with_diffusion_and_step_n(oracle, cuda_lib):
    P.check_steps(oracle, cuda_lib, 333, 6, 3, P.random_rects(333, 8, 2), with_impulses=True,
                  diffusion=1e-4, use_step_n=True)


def test_clone_default_reset(oracle, cuda_lib):
    from equilibrium_b200 import Fluid, Rectangle
    dev = Fluid.default(lib_path=cuda_lib)
    ref = oracle.RefFluid(128, 0.02, 16)
    ref.init()
    P.assert_state_equal(dev, ref, "Fluid::default")
    dev.fill_obstacle(Rectangle.default())
    ref.fill_rect(80, 80, 110, 110)
    assert int(dev.cells_type.sum()) == 1408
    twin = dev.clone()
    dev.step()
    ref.step()
    P.assert_state_equal(dev, ref, "after clone + step")
    twin.step()
    P.assert_state_equal(twin, ref, "the clone steps to the same state")
    dev.reset_walls()
    assert int(dev.cells_type.sum()) == 2 * (128 + 126)


def test_config3_4096_single_ops_reduced_k(oracle, cuda_lib):
    # BASELINE config 3 geometry (4096^2, 64 random rectangles, seed 4096); K reduced so the
    # oracle finishes in seconds.  Full K=40 is covered by the property tests below.
    n = 4096
    rects = P.random_rects(n, 64, 4096)
    P.check_lin_solve(oracle, cuda_lib, n, 3, rects, P.COL)
    P.check_lin_solve(oracle, cuda_lib, n, 2, rects, P.ROW)
    P.check_advect(oracle, cuda_lib, n, rects, P.ROW, vscale=2.0)
    P.check_project(oracle, cuda_lib, n, 2, rects)


def test_config3_4096_one_step_k4(oracle, cuda_lib):
    P.check_steps(oracle, cuda_lib, 4096, 4, 1, P.random_rects(4096, 64, 4096), with_impulses=True)


def test_config3_4096_k40_linearity_and_determinism(cuda_lib):
    """Size-independent properties at the full config-3 size and K=40:
    (1) lin_solve is exactly linear under scaling by a power of two (scaling by 4 is
        exact in binary floating point), every cell bit-compared;
    (2) two runs give identical bits (the wavefront schedule never leaks into results)."""
    from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs
    n, k = 4096, 40
    rng = np.random.default_rng(0)
    dev = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=cuda_lib)
    for r in P.random_rects(n, 64, 4096):
        dev.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    outs = []
    for scale in (1.0, 4.0, 1.0):
        dev.upload("velocities_x", x * np.float32(scale))
        dev.upload("velocities_x0", x0 * np.float32(scale))
        dev.op_lin_solve(P.COL, "velocities_x", "velocities_x0", 0.37, 2.48, k)
        outs.append(dev.download("velocities_x"))
    assert P.bits_equal(outs[0], outs[2]), "exact mode is not deterministic"
    assert P.bits_equal(outs[0] * np.float32(4.0), outs[1]), "lin_solve(4x, 4x0) != 4 lin_solve(x, x0)"


def test_config4_16384_lin_solve_direct(oracle, cuda_lib):
    # BASELINE config 4 geometry (16384^2, 16 rectangles, seed 16384), one AdjustColumn
    # lin_solve with K=2 compared directly (about 10 s of oracle time)
    n = 16384
    P.check_lin_solve(oracle, cuda_lib, n, 2, P.random_rects(n, 16, 16384), P.COL)


@pytest.mark.parametrize("n", [128, 1001])
def test_render_rgba_and_snapshots(oracle, cuda_lib, n):
    P.check_render_and_snapshot(oracle, cuda_lib, n, P.random_rects(n, 6, 3))


def test_render_rgba_golden_pixels(cuda_lib):
    P.check_golden_render(cuda_lib)

