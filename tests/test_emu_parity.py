"""CPU-only: the PRODUCT kernel sources (equilibrium_b200/csrc) compiled against the
host SIMT emulator in tests/emu and driven through the same C ABI, bit-compared
with the oracle.  This checks kernel logic, the wavefront's ticket/flag protocol
under real thread concurrency and the host-side composition of step() without a
GPU; the -m gpu tests repeat the comparisons on the real device."""
import numpy as np
import pytest

import parity as P

RECTS64 = [(20, 30, 40, 45), (5, 5, 9, 60), (31, 1, 34, 33), (50, 50, 63, 63)]


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_set_boundaries(oracle, emu_lib, orient):
    P.check_set_boundaries(oracle, emu_lib, 64, RECTS64, orient)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,k,rects", [(64, 3, RECTS64), (45, 2, [(1, 1, 44, 2), (10, 3, 11, 44)]), (33, 1, [])])
def test_lin_solve_exact(oracle, emu_lib, orient, n, k, rects):
    P.check_lin_solve(oracle, emu_lib, n, k, rects, orient)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,k", [(97, 3), (98, 4), (99, 5), (100, 1), (130, 3), (131, 2)])
def test_lin_solve_exact_last_band_under_skew(oracle, emu_lib, orient, n, k):
    # Temporal blocking moves a band up two rows per fused iteration, so the row N-2 (and the
    # frame row behind it) is owned by different bands at different sub-steps around N = 32m+2;
    # odd K leaves a shorter last group.
    P.check_lin_solve(oracle, emu_lib, n, k, [(n - 12, 3, n - 1, 9), (5, n - 6, 40, n - 1), (30, 30, 36, 66)], orient)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_lin_solve_exact_obstacles_touching_the_frame(oracle, emu_lib, orient):
    # Rows 1 / N-2 and columns 1 / N-2 normally hold plain frame-adjacent fix-ups, which the fast loops take
    # in their stride; obstacles glued to each side of the frame break that pattern for some chunks only.
    n = 200
    rects = [(50, 1, 70, 4), (120, n - 5, 150, n - 1), (1, 90, 6, 120), (n - 4, 20, n - 1, 60), (100, 100, 104, 104)]
    P.check_lin_solve(oracle, emu_lib, n, 4, rects, orient)


@pytest.mark.parametrize("n,k", [(98, 4), (131, 3)])
def test_lin_solve_passive_with_a_walled_off_column(oracle, emu_lib, n, k):
    # An obstacle spanning the full height leaves columns without a NoWall cell: their frame-row
    # cells must keep their values (quirk Q6), so the Passive frame bands fall back to the general loop.
    P.check_lin_solve(oracle, emu_lib, n, k, [(40, 1, 43, n - 1), (5, n - 6, 30, n - 1)], P.PASSIVE)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_lin_solve_exact_fast_and_general_macro_steps(oracle, emu_lib, orient):
    # N >= 97 is needed for a macro step with every lane on interior columns; with one small
    # rectangle most (band, chunk) pairs are code-free and take the branch-free fast loop,
    # the ones around the rectangle, the frame columns and the ragged last band do not.
    P.check_lin_solve(oracle, emu_lib, 168, 2, [(100, 70, 108, 101)], orient)


A0_POISONS = [None, ("x0", 20, 30, -0.0), ("x0", 5, 5, float("nan")), ("x", 0, 7, float("inf")), ("x", 40, 41, -3e38),
              ("x0", 1, 1, 2e37), ("x", 63, 63, float("nan"))]


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("poison", A0_POISONS)
def test_lin_solve_zero_coefficient_shortcut_and_its_guard(oracle, emu_lib, orient, poison):
    # the reference's default diffusion is 0 (configs.rs:50-60): a = 0, c = 1
    P.check_lin_solve_a0(oracle, emu_lib, 64, 3, RECTS64, orient, poison)


def test_lin_solve_zero_coefficient_red_black(oracle, emu_lib):
    P.check_lin_solve_a0(oracle, emu_lib, 70, 5, [(10, 10, 30, 20)], P.PASSIVE, None, mode="red_black")
    P.check_lin_solve_a0(oracle, emu_lib, 70, 5, [(10, 10, 30, 20)], P.PASSIVE, ("x0", 33, 21, -0.0), mode="red_black")


def test_lin_solve_zero_iterations_is_a_no_op(oracle, emu_lib):
    P.check_lin_solve(oracle, emu_lib, 32, 0, [], P.ROW)


def test_project(oracle, emu_lib):
    P.check_project(oracle, emu_lib, 64, 3, RECTS64)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_advect_with_row_break(oracle, emu_lib, orient):
    P.check_advect(oracle, emu_lib, 48, [(10, 10, 20, 30)], orient, vscale=4.0)


def test_full_steps_with_impulses(oracle, emu_lib):
    P.check_steps(oracle, emu_lib, 64, 4, 3, [(20, 30, 40, 45)], with_impulses=True)


def test_full_steps_ragged_size_and_diffusion(oracle, emu_lib):
    P.check_steps(oracle, emu_lib, 50, 2, 2, [(10, 10, 20, 45), (30, 1, 31, 49)], with_impulses=True,
                  diffusion=1e-3)


def test_step_n_with_sources_and_clone(oracle, emu_lib):
    dev, ref = P.check_steps(oracle, emu_lib, 40, 2, 3, [(5, 5, 15, 15)], with_impulses=True, use_step_n=True)
    twin = dev.clone()
    dev.step()
    ref.step()
    P.assert_state_equal(dev, ref, "after clone + step")
    assert not P.bits_equal(twin.download("velocities_x"), dev.download("velocities_x"))


def test_default_double_init_and_reset_walls(oracle, emu_lib):
    from equilibrium_b200 import Fluid, Rectangle
    dev = Fluid.default(lib_path=emu_lib)
    ref = oracle.RefFluid(128, 0.02, 16)
    ref.init()
    P.assert_state_equal(dev, ref, "Fluid::default")
    dev.fill_obstacle(Rectangle.default())
    assert int(dev.cells_type.sum()) == 1408          # renderer_helpers.rs:222-252
    dev.reset_walls()
    assert int(dev.cells_type.sum()) == 2 * (128 + 126)


def test_fill_obstacle_clamps_like_idx(oracle, emu_lib):
    dev, ref = P.make_pair(oracle, emu_lib, 32, 1)
    for r in [(-5, 10, 4, 12), (28, 28, 40, 40), (3, 3, 3, 9), (9, 9, 5, 12)]:
        class R:
            def get_approximate_points(self, r=r):
                return [(r[0], r[1]), (r[2], r[3])]
        dev.fill_obstacle(R())
        ref.fill_rect(*r)
    assert np.array_equal(dev.cells_type, ref.cells)


def test_upload_rejects_open_frame(emu_lib):
    from equilibrium_b200 import Fluid, FluidConfigs, SimulationConfigs, EquilibriumError
    dev = Fluid(FluidConfigs(), SimulationConfigs(0.02, 1, 24), lib_path=emu_lib)
    cells = dev.cells_type
    cells[0, 5] = 0
    with pytest.raises(EquilibriumError):
        dev.upload("cells_type", cells)


def test_red_black_close_to_oracle_red_black(oracle, emu_lib):
    # the fast path is checked against the red-black restatement bitwise and against the
    # lexicographic oracle within tolerance (tests/test_red_black.py does the latter on GPU)
    rng = np.random.default_rng(0)
    n, k = 48, 6
    dev, ref = P.make_pair(oracle, emu_lib, n, k, [(10, 10, 20, 30)], mode="red_black")
    for orient in (P.ROW, P.COL, P.PASSIVE):
        x, x0 = P.rnd(rng, n), P.rnd(rng, n)
        dev.upload("velocities_x", x)
        dev.upload("velocities_x0", x0)
        dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
        oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
        got = dev.download("velocities_x")
        assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("n", [200, 301])
@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_red_black_tiled_across_tiles(oracle, emu_lib, orient, n):
    # k_rb_reg tiles write 232 x 40 cells: 200 spans several tile rows, 301 (odd: the last column pair is half
    # outside the grid) also two tile columns, with rectangles across the tile seams; 5 iterations = one full
    # pass of 4 plus a tail of 1
    rng = np.random.default_rng(1)
    k = 5
    rects = [(60, 50, 140, 70), (120, 100, 131, 190), (1, 128, 40, 129)]
    if n > 232:
        rects += [(220, 30, 250, 50), (225, 200, 240, 299), (231, 90, 233, 92)]
    dev, ref = P.make_pair(oracle, emu_lib, n, k, rects, mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("n", [24, 101, 128, 160])
@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_red_black_single_cta_kernel(oracle, emu_lib, orient, n):
    # grids up to N * P = 32768 cells (the reference's 128^2 default scene) run all K iterations in one launch of one CTA
    # (k_rb_small): 128 is the 8-rows-per-thread instantiation at its limit, 160 the 16-row one, 101 an odd size with
    # idle threads in the last row group
    rng = np.random.default_rng(n)
    k = 7
    rects = [(3, 3, 9, 8)] if n < 40 else [(10, 10, 20, 30), (n - 30, 5, n - 12, n - 8), (40, n - 4, 60, n - 2)]
    dev, ref = P.make_pair(oracle, emu_lib, n, k, rects, mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("kernel", ["slide", "stream"])
@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_red_black_sliding_window_kernels(oracle, emu_lib, orient, kernel, monkeypatch):
    # k_rb_slide (2 iterations per pass) and k_rb_stream (4 per pass, rows brought in by bulk copies) are the large-grid
    # kernels; EQ_RB_KERNEL forces them on a grid small enough for the emulator.  420 columns = 4 / 5 strips, the
    # rectangles sit in one corner so that interior tasks without mirror codes (no range tests, no set_boundaries),
    # interior tasks with codes and edge tasks all occur; 11 iterations = two passes of 4, one of 2, one of 1
    monkeypatch.setenv("EQ_RB_KERNEL", kernel)
    monkeypatch.setenv("EQ_RQ_SEGS", "4")
    rng = np.random.default_rng(7)
    n, k = 420, 11
    rects = [(300, 40, 330, 90), (310, 300, 318, 380), (2, 200, 9, 203)]
    dev, ref = P.make_pair(oracle, emu_lib, n, k, rects, mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


def test_red_black_default_kernel_choice_at_2048(oracle, emu_lib):
    # no EQ_RB_KERNEL: a 2048-column grid takes k_rb_stream by default, with the task shapes the library picks for a
    # launch of less than one wave (96-row segments, wall-strip tasks sized to match); 5 iterations = one pass of 4 through
    # k_rb_stream and one of 1 through k_rb_slide.  ~20 s under the emulator.
    rng = np.random.default_rng(1)
    n, k = 2048, 5
    dev, ref = P.make_pair(oracle, emu_lib, n, k, [(300, 40, 330, 900), (1000, 1200, 1900, 1230)], mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(P.ROW, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(P.ROW, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("n,steps", [(64, 1), (101, 0)])      # (the GPU suite runs more frames; emulated steps are slow)
def test_render_rgba_and_snapshots(oracle, emu_lib, n, steps):
    P.check_render_and_snapshot(oracle, emu_lib, n, [(10, 10, 20, 30), (40, 5, 50, 60)], steps=steps)


def test_render_rgba_golden_pixels(emu_lib):
    P.check_golden_render(emu_lib)


def test_current_simulation_loop(oracle, emu_lib):
    P.check_current_simulation(oracle, emu_lib)
