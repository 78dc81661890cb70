"""CPU-only: row-slab decomposition with several handles in one process, each
driven by its own thread, on the emulated build of the product sources.  The ranks exchange
halo rows, solver flags and raw rows through each other's memory exactly as the GPU build
does over NVLink; the assembled slabs must be bit-identical to the single-domain oracle."""
import threading

import numpy as np
import pytest

import parity as P
from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs


def run_ranks(world, body):
    """body(rank, barrier) in one thread per rank; re-raises the first failure."""
    barrier = threading.Barrier(world)
    errors, results = [], [None] * world

    def wrap(r):
        try:
            results[r] = body(r, barrier)
        except Exception as e:   # noqa: BLE001
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=wrap, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def make_rank_fluids(emu_lib, world, n, k, rects, mode="exact", diffusion=0.0):
    fluids = [Fluid(FluidConfigs(diffusion=diffusion), SimulationConfigs(0.02, k, n), lib_path=emu_lib,
                    mode=mode, rank=r, world=world) for r in range(world)]
    blobs = [f.ipc_blob() for f in fluids]
    for f in fluids:
        f.ipc_attach(blobs)
        for (x0, y0, x1, y1) in rects:
            f.fill_obstacle(Rectangle((x0, y0), (x1, y1), n))
    return fluids


def assemble(fluids, name):
    n = fluids[0].simulation_configs.size
    out = np.zeros((n, n), dtype=np.float32)
    covered = np.zeros(n, dtype=bool)
    for f in fluids:
        r0, rows = f.download_owned(name)
        out[r0:r0 + rows.shape[0]] = rows
        covered[r0:r0 + rows.shape[0]] = True
    assert covered.all()
    return out


@pytest.mark.parametrize("plan", ["slab", "replica"])
@pytest.mark.parametrize("world,n,k", [(2, 96, 3), (3, 130, 2)])
def test_steps_match_single_domain_oracle(oracle, emu_lib, world, n, k, plan, monkeypatch):
    monkeypatch.setenv("EQ_EXACT_PLAN", plan)
    rects = [(20, 28, 40, 40), (50, 60, 70, 66), (5, 30, 9, 90)]
    fluids = make_rank_fluids(emu_lib, world, n, k, rects)
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    slabs = [f.owned_rows() for f in fluids]
    assert slabs[0][0] == 0 and slabs[-1][1] == n
    assert all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))
    imp = P.impulses(n, 2, 7)

    def body(r, barrier):
        f = fluids[r]
        for (_, x, y, ax, ay) in imp:
            f.add_velocity(x, y, ax, ay)
            f.step()
        f.sync()

    for (_, x, y, ax, ay) in imp:
        ref.add_velocity(x, y, ax, ay)
        ref.step()
    run_ranks(world, body)
    for name, fid in P.F32_FIELDS:
        got, want = assemble(fluids, name), ref.field(fid)
        assert P.bits_equal(got, want), f"{name}: {P.describe_diff(got, want)}"


@pytest.mark.parametrize("plan", ["slab", "replica"])
@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
def test_lin_solve_across_two_slabs(oracle, emu_lib, orient, plan, monkeypatch):
    # exact mode on row slabs has two plans (eq_api.cu lin_solve_dispatch): the row-slab wavefront kernel, or every rank
    # gathering the other slabs over peer memory and running the single-GPU kernel on the whole grid (the default up to 4
    # ranks); each rank only holds its own rows when the solve starts
    monkeypatch.setenv("EQ_EXACT_PLAN", plan)
    n, k, world = 100, 4, 2
    rects = [(10, 40, 60, 70), (70, 10, 80, 90)]    # straddle the slab boundary (row 65)
    fluids = make_rank_fluids(emu_lib, world, n, k, rects)
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    rng = np.random.default_rng(3)
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    for f in fluids:
        f.upload("velocities_x", x)
        f.upload("velocities_x0", x0)

    def body(r, barrier):
        fluids[r].op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
        fluids[r].sync()

    run_ranks(world, body)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells)
    got = assemble(fluids, "velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("world,n,k", [(2, 96, 3), (2, 200, 6), (3, 200, 9)])
def test_red_black_across_slabs_matches_its_restatement(oracle, emu_lib, world, n, k):
    # k > 4: the launches after the first read ghost rows that the neighbours' k_rb_reg pushed itself (3 ranks: the
    # middle one pushes both ways)
    rects = [(10, 40, 60, 70)]
    fluids = make_rank_fluids(emu_lib, world, n, k, rects, mode="red_black")
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    rng = np.random.default_rng(4)
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    for f in fluids:
        f.upload("velocities_x", x)
        f.upload("velocities_x0", x0)

    def body(r, barrier):
        fluids[r].op_lin_solve(P.COL, "velocities_x", "velocities_x0", 0.37, 2.48, k)
        fluids[r].sync()

    run_ranks(world, body)
    oracle.lin_solve(P.COL, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = assemble(fluids, "velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("kernel", ["stream", "slide"])
@pytest.mark.parametrize("world,n,k", [(2, 200, 9), (3, 260, 6)])
def test_red_black_large_grid_kernels_across_slabs(oracle, emu_lib, world, n, k, kernel, monkeypatch):
    # the sliding-window kernels (the default from 2048 columns on) on row slabs: every pass starts with the exchange of
    # 12 ghost rows, segments at a slab edge recompute the neighbour's rows
    monkeypatch.setenv("EQ_RB_KERNEL", kernel)
    monkeypatch.setenv("EQ_RQ_SEGS", "2")
    rects = [(10, 40, 60, 70), (120, 90, 130, 120)]
    fluids = make_rank_fluids(emu_lib, world, n, k, rects, mode="red_black")
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    rng = np.random.default_rng(11)
    for orient in (P.ROW, P.PASSIVE):
        x, x0 = P.rnd(rng, n), P.rnd(rng, n)
        for f in fluids:
            f.upload("velocities_x", x)
            f.upload("velocities_x0", x0)

        def body(r, barrier):
            fluids[r].op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
            fluids[r].sync()

        run_ranks(world, body)
        oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
        got = assemble(fluids, "velocities_x")
        assert P.bits_equal(got, x), P.describe_diff(got, x)


def test_unattached_handle_refuses_to_step(emu_lib):
    from equilibrium_b200 import EquilibriumError
    f = Fluid(FluidConfigs(), SimulationConfigs(0.02, 1, 96), lib_path=emu_lib, rank=0, world=2)
    with pytest.raises(EquilibriumError):
        f.step()


def test_device_noise_across_two_slabs(oracle, emu_lib):
    # every rank draws the same Philox impulse for the centre cell; only the owner's copy of that row is authoritative
    world, n, k, seed = 2, 96, 2, 99
    rects = [(20, 28, 40, 40)]
    fluids = make_rank_fluids(emu_lib, world, n, k, rects)
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    nz = fluids[0].device_noise(seed)

    def body(r, barrier):
        fluids[r].step_n_noise(2, seed)
        fluids[r].sync()

    for fr in range(2):
        ref.add_velocity(*oracle.noise_impulse(seed, fr, n, nz.cos_t, nz.sin_t, nz.gain))
        ref.step()
    run_ranks(world, body)
    for name, fid in P.F32_FIELDS:
        got, want = assemble(fluids, name), ref.field(fid)
        assert P.bits_equal(got, want), f"{name}: {P.describe_diff(got, want)}"


@pytest.mark.parametrize("mode", ["exact", "red_black"])
@pytest.mark.parametrize("poison_row", [None, 70])
def test_zero_coefficient_shortcut_across_slabs(oracle, emu_lib, mode, poison_row):
    """lin_solve with a = 0, c = 1 on three slabs: the guard of the shortcut is OR-ed over the ranks, so a -0.0 in ONE
    rank's rows (row 70 is the last slab's) must send every rank through the sweeps; without it every rank copies."""
    world, n, k = 3, 96, 4
    rects = [(20, 28, 40, 40), (50, 60, 70, 66)]
    fluids = make_rank_fluids(emu_lib, world, n, k, rects, mode=mode)
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    rng = np.random.default_rng(3)
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    x0[rng.random((n, n)) < 0.1] = 0.0
    if poison_row is not None:
        x0[poison_row, 33] = np.float32(-0.0)
    for f in fluids:
        f.upload("velocities_x", x)
        f.upload("velocities_x0", x0)

    def body(r, barrier):
        fluids[r].op_lin_solve(P.PASSIVE, "velocities_x", "velocities_x0", 0.0, 1.0, k)
        fluids[r].sync()

    run_ranks(world, body)
    oracle.lin_solve(P.PASSIVE, x, x0, 0.0, 1.0, k, ref.cells, red_black=(mode == "red_black"))
    got = assemble(fluids, "velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)
