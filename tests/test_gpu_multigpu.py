"""GPU tests of the row-slab path: several handles in this process, one per device,
each driven by its own thread; neighbours are mapped with peer access and every kernel reads /
writes them over NVLink.  Skipped on a one-GPU box (bench.py --gpus N covers the one-process-
per-GPU / CUDA-IPC flavour of the same code)."""
import numpy as np
import pytest

import parity as P
from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs, _lib
from test_emu_multirank import assemble, run_ranks

pytestmark = pytest.mark.gpu


def ndev(cuda_lib):
    return _lib.load(cuda_lib).eq_device_count()


def make(cuda_lib, world, n, k, rects, mode="exact"):
    fluids = [Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=cuda_lib, mode=mode,
                    device=r, rank=r, world=world) for r in range(world)]
    blobs = [f.ipc_blob() for f in fluids]
    for f in fluids:
        f.ipc_attach(blobs)
        for (x0, y0, x1, y1) in rects:
            f.fill_obstacle(Rectangle((x0, y0), (x1, y1), n))
    return fluids


@pytest.mark.parametrize("world,n,k,frames", [(2, 512, 8, 3), (4, 1024, 20, 2), (8, 1024, 5, 2)])
def test_slabs_match_single_domain_oracle(oracle, cuda_lib, world, n, k, frames):
    if ndev(cuda_lib) < world:
        pytest.skip(f"needs {world} GPUs")
    rects = P.random_rects(n, 12, n + world)
    fluids = make(cuda_lib, world, n, k, rects)
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    imp = P.impulses(n, frames, 5)

    def body(r, barrier):
        fluids[r].step_n(frames, imp)
        fluids[r].sync()

    for (_, x, y, ax, ay) in imp:
        ref.add_velocity(x, y, ax, ay)
        ref.step()
    run_ranks(world, body)
    for name, fid in P.F32_FIELDS:
        got, want = assemble(fluids, name), ref.field(fid)
        assert P.bits_equal(got, want), f"{name}: {P.describe_diff(got, want)}"


def test_two_slabs_equal_one_gpu_at_4096(cuda_lib):
    """Size-independent property at a BASELINE size: the decomposition does not change a bit."""
    if ndev(cuda_lib) < 2:
        pytest.skip("needs 2 GPUs")
    n, k = 4096, 10
    rects = P.random_rects(n, 64, 4096)
    one = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=cuda_lib)
    for r in rects:
        one.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
    two = make(cuda_lib, 2, n, k, rects)
    imp = P.impulses(n, 2, 9)
    one.step_n(2, imp)

    def body(r, barrier):
        two[r].step_n(2, imp)
        two[r].sync()

    run_ranks(2, body)
    for name, _ in P.F32_FIELDS:
        assert P.bits_equal(assemble(two, name), one.download(name)), name


@pytest.mark.parametrize("world,n,k", [(2, 1024, 9), (8, 2048, 12)])
def test_red_black_slabs_match_the_restatement(oracle, cuda_lib, world, n, k):
    """k_rb_reg over row slabs: the first launch reads ghost rows copied by k_halo_exchange, the later ones ghost rows
    the neighbours' k_rb_reg pushed over NVLink itself.  Bit-identical to the one-domain red-black restatement."""
    if ndev(cuda_lib) < world:
        pytest.skip(f"needs {world} GPUs")
    rects = P.random_rects(n, 10, n + 3 * world)
    fluids = make(cuda_lib, world, n, k, rects, mode="red_black")
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    rng = np.random.default_rng(world)
    for orient in (P.ROW, P.COL, P.PASSIVE):
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
        assert P.bits_equal(got, x), f"orient {orient}: {P.describe_diff(got, x)}"
