"""Red-black fast path (EQ_MODE_RED_BLACK): same formula and iteration count as
lin_solve (fluid.rs:301-325), cells with (i+j) even first, then odd, then
set_boundaries.

Stated tolerance:
  * bit-identical to the red-black restatement in the oracle (ref_lin_solve_red_black);
  * against the reference's lexicographic order, on the smooth default-style scene (no
    impulses) after 4 frames: velocity fields within 5e-2 relative L2, density within
    1e-1 relative L2 (a sharp-edged blob: small transport differences show), and the L2 norm of the velocity divergence (the residual the
    projection is there to shrink) within 10 % of the oracle's.
  With the scripted +-2N impulses the flow is chaotic and the two orderings separate after a
  frame or two (each is K sweeps away from the converged solve), so no field tolerance is
  claimed there -- only that both stay finite and the divergence residual stays comparable.
"""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


def div_l2(vx, vy, n):
    vx, vy = vx.astype(np.float64), vy.astype(np.float64)
    d = -0.5 * ((vx[1:-1, 2:] - vx[1:-1, :-2]) + (vy[2:, 1:-1] - vy[:-2, 1:-1])) / n
    return float(np.linalg.norm(d))


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,k", [(128, 16), (1000, 5)])
def test_bitwise_against_red_black_restatement(oracle, cuda_lib, orient, n, k):
    rng = np.random.default_rng(n)
    dev, ref = P.make_pair(oracle, cuda_lib, n, k, P.random_rects(n, 6, 1), mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("orient", [P.ROW, P.COL, P.PASSIVE])
@pytest.mark.parametrize("n,k", [(2500, 11), (3001, 8)])
def test_bitwise_large_grid_kernels(oracle, cuda_lib, orient, n, k):
    # grids of 2048 columns or more take k_rb_stream (4 iterations per pass, k_rb_slide for what is left of K): 25 / 29
    # strips of 104 columns, several segments, rectangles across the strip and segment seams; 3001 is odd (the last
    # column quad is half outside the grid)
    rng = np.random.default_rng(n)
    dev, ref = P.make_pair(oracle, cuda_lib, n, k, P.random_rects(n, 24, 2), mode="red_black")
    x, x0 = P.rnd(rng, n), P.rnd(rng, n)
    dev.upload("velocities_x", x)
    dev.upload("velocities_x0", x0)
    dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
    oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
    got = dev.download("velocities_x")
    assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("kernel", ["stream", "slide", "reg"])
def test_bitwise_every_kernel_on_one_grid(oracle, cuda_lib, kernel, monkeypatch):
    # the three red-black kernels are interchangeable: EQ_RB_KERNEL forces each of them on the same 1500^2 problem
    monkeypatch.setenv("EQ_RB_KERNEL", kernel)
    n, k = 1500, 7
    rng = np.random.default_rng(5)
    dev, ref = P.make_pair(oracle, cuda_lib, n, k, P.random_rects(n, 12, 3), mode="red_black")
    for orient in (P.ROW, P.COL, P.PASSIVE):
        x, x0 = P.rnd(rng, n), P.rnd(rng, n)
        dev.upload("velocities_x", x)
        dev.upload("velocities_x0", x0)
        dev.op_lin_solve(orient, "velocities_x", "velocities_x0", 0.37, 2.48, k)
        oracle.lin_solve(orient, x, x0, 0.37, 2.48, k, ref.cells, red_black=True)
        got = dev.download("velocities_x")
        assert P.bits_equal(got, x), P.describe_diff(got, x)


@pytest.mark.parametrize("n,k", [(128, 16), (512, 20)])
def test_tolerance_against_lexicographic_reference(oracle, cuda_lib, n, k):
    dev, ref = P.make_pair(oracle, cuda_lib, n, k, P.random_rects(n, 4, n), mode="red_black")
    for _ in range(4):
        dev.step()
        ref.step()
    vx, vy, d = dev.download("velocities_x"), dev.download("velocities_y"), dev.download("density")
    assert rel_l2(vx, ref.vx) <= 5e-2
    assert rel_l2(vy, ref.vy) <= 5e-2
    assert rel_l2(d, ref.density) <= 1e-1
    dd, dr = div_l2(vx, vy, n), div_l2(ref.vx, ref.vy, n)
    assert abs(dd - dr) <= 0.10 * dr, (dd, dr)


def test_impulses_stay_finite_and_residual_comparable(oracle, cuda_lib):
    n, k = 256, 20
    dev, ref = P.make_pair(oracle, cuda_lib, n, k, P.random_rects(n, 4, n), mode="red_black")
    for (_, x, y, ax, ay) in P.impulses(n, 4, 0):
        dev.add_velocity(x, y, ax, ay)
        ref.add_velocity(x, y, ax, ay)
        dev.step()
        ref.step()
    vx, vy = dev.download("velocities_x"), dev.download("velocities_y")
    assert np.isfinite(vx).all() and np.isfinite(vy).all() and np.isfinite(dev.download("density")).all()
    dd, dr = div_l2(vx, vy, n), div_l2(ref.vx, ref.vy, n)
    assert 0.3 * dr <= dd <= 3.0 * dr, (dd, dr)


def _metrics(rb, ex, n):
    """Field rel-L2 of red-black against exact (chunked: the 16384^2 fields are 1 GiB each) and the divergence residuals."""
    out = {}
    for name in ("velocities_x", "velocities_y", "density"):
        a, b = rb.download(name), ex.download(name)
        num = den = 0.0
        for r0 in range(0, n, 1024):
            da = a[r0:r0 + 1024].astype(np.float64)
            db = b[r0:r0 + 1024].astype(np.float64)
            num += float(((da - db) ** 2).sum())
            den += float((db ** 2).sum())
        out[name] = (num / max(den, 1e-300)) ** 0.5
        assert np.isfinite(a).all()
    out["div_rb"] = rb.divergence_l2("velocities_x", "velocities_y")
    out["div_exact"] = ex.divergence_l2("velocities_x", "velocities_y")
    return out


@pytest.mark.parametrize("cfg,n,k,nrect,frames", [("c3", 4096, 40, 64, (1, 4)), ("c4", 16384, 20, 16, (1, 4))])
def test_tolerance_against_gpu_exact_at_baseline_configs(cuda_lib, cfg, n, k, nrect, frames):
    """BASELINE configs 3 and 4 ("bit-order wavefront vs red-black modes"): the exact mode -- bit-identical to the oracle at
    these sizes (test_gpu_parity.py) -- is the reference here (GPU-exact as the secondary oracle where the
    CPU oracle takes minutes per frame).  Same scene, same K.  The stated tolerance of the fast path."""
    import json
    import os
    from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs
    rects = P.random_rects(n, nrect, n)
    fl = {}
    for mode in ("exact", "red_black"):
        f = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=cuda_lib, mode=mode)
        for (x0, y0, x1, y1) in rects:
            f.fill_obstacle(Rectangle((x0, y0), (x1, y1), n))
        fl[mode] = f
    done, report = 0, {}
    for target in frames:
        for f in fl.values():
            f.step_n(target - done)
        done = target
        m = _metrics(fl["red_black"], fl["exact"], n)
        report[f"frames_{target}"] = m
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/rb_tolerance_{cfg}.json", "w") as fh:
            json.dump(report, fh, indent=1)
    except OSError:
        pass
    for key, m in report.items():
        # stated tolerance: velocity within 2e-2 relative L2 of the exact mode after one frame and within 1e-1
        # after four (two orderings of K un-converged sweeps drift apart frame by frame), density within 5e-2, and the
        # divergence residual the projection leaves within 10 % of the exact mode's
        vtol = 2e-2 if key == "frames_1" else 1e-1
        assert m["velocities_x"] <= vtol and m["velocities_y"] <= vtol, (key, m)
        assert m["density"] <= 5e-2, (key, m)
        assert abs(m["div_rb"] - m["div_exact"]) <= 0.10 * m["div_exact"], (key, m)
