"""device-side sources.  The reference's add_noise (fluid.rs:575-599) draws from an unseeded
thread_rng, so no sequence of it can be reproduced; the device path keeps its structure (random grid point, rotated
about the centre, twice the result added to the centre cell's velocity) with a Philox4x32-10 stream.  Philox is pinned
by the Random123 known-answer vectors; the impulse by two independent restatements (C and numpy) that must agree bit
for bit; the kernels by the oracle on the emulated build here and on the GPU under -m gpu."""
import numpy as np
import pytest

import parity as P
from equilibrium_b200 import EquilibriumError, Fluid, FluidConfigs, SimulationConfigs


class _LazyPyref:
    """oracle.pyref pulls numba in; keep that out of collection (a `-m gpu` run never needs it at import time)."""
    def __getattr__(self, name):
        from oracle import pyref as m
        return getattr(m, name)


pyref = _LazyPyref()

# Random123 kat_vectors, philox4x32 with 10 rounds: counter, key, expected
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(oracle, ctr, key, want):
    assert oracle.philox4x32_10(ctr, key) == want
    assert pyref.philox4x32_10(ctr, key) == want


def test_noise_impulse_restatements_agree(oracle):
    rng = np.random.default_rng(0)
    for _ in range(300):
        seed, frame = int(rng.integers(0, 2**63)) * 2 + 1, int(rng.integers(0, 2**40))
        n = int(rng.integers(20, 40000))
        th = float(rng.uniform(-7, 7))
        cs, sn = float(np.float32(np.cos(th))), float(np.float32(np.sin(th)))
        a, b = oracle.noise_impulse(seed, frame, n, cs, sn), pyref.noise_impulse(seed, frame, n, cs, sn)
        assert a[:2] == b[:2] == (n // 2, n // 2)
        assert np.float32(a[2]).tobytes() == np.float32(b[2]).tobytes()
        assert np.float32(a[3]).tobytes() == np.float32(b[3]).tobytes()


def test_noise_points_cover_the_grid(oracle):
    # identity rotation, gain 1: the impulse IS the random point; it must stay inside [0, N) and reach both ends
    n = 64
    pts = np.array([oracle.noise_impulse(7, fr, n, 1.0, 0.0, 1.0)[2:] for fr in range(4000)])
    assert pts.min() == 0 and pts.max() == n - 1
    assert np.all(pts == np.floor(pts))
    assert abs(pts.mean() - (n - 1) / 2) < 1.0
    # another seed / another frame gives another stream
    assert oracle.noise_impulse(7, 0, n, 1.0, 0.0, 1.0) != oracle.noise_impulse(8, 0, n, 1.0, 0.0, 1.0)


# (the golden run below also cuts one stream into two calls)
@pytest.mark.parametrize("n,k,frames,seed,first", [(97, 2, 3, 0xDEADBEEFCAFE, 2**33 + 5)])
def test_device_noise_emulated(oracle, emu_lib, n, k, frames, seed, first):
    P.check_device_noise(oracle, emu_lib, n, k, frames, [(10, 12, 30, 20)], seed, first)


def test_add_noise_device_equals_the_fused_loop(emu_lib):
    # { add_noise; step } x 2 through eq_add_noise == eq_step_n_noise(2)
    n, k = 64, 2
    a = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=emu_lib)
    b = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=emu_lib)
    a.step_n_noise(2, 42, first_frame=7)
    for fr in (7, 8):
        b.add_noise_device(42, fr)
        b.step()
    for name, _ in P.F32_FIELDS:
        assert P.bits_equal(a.download(name), b.download(name)), name


@pytest.mark.parametrize("n", [64, 97, 130])
def test_add_source_emulated(oracle, emu_lib, n):
    P.check_add_source(oracle, emu_lib, n)


def test_add_source_rejects_aliased_fields(emu_lib):
    f = Fluid(FluidConfigs(), SimulationConfigs(0.02, 1, 64), lib_path=emu_lib)
    with pytest.raises(EquilibriumError):
        f.op_add_source("density", "density", 1.0)
    with pytest.raises(EquilibriumError):
        f.op_add_source("density", "cells_type", 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,frames,seed,first", [(128, 16, 3, 1, 0), (1001, 4, 2, 0xDEADBEEFCAFE, 2**33 + 5)])
def test_device_noise_gpu(oracle, cuda_lib, n, k, frames, seed, first):
    P.check_device_noise(oracle, cuda_lib, n, k, frames, P.random_rects(n, 5, n), seed, first)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [128, 1001, 4096])
def test_add_source_gpu(oracle, cuda_lib, n):
    P.check_add_source(oracle, cuda_lib, n)


def test_oracle_matches_golden_noise_impulses(oracle):
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "device_noise.json")) as f:
        g = json.load(f)
    cs, sn = float.fromhex(g["cos_t"]), float.fromhex(g["sin_t"])
    for fr, (x, y, ax, ay) in enumerate(g["impulses"]):
        for impl in (oracle.noise_impulse, pyref.noise_impulse):
            got = impl(g["seed"], g["first_frame"] + fr, g["n"], cs, sn, g["gain"])
            assert got == (x, y, float.fromhex(ax), float.fromhex(ay)), (fr, impl.__module__)


def test_golden_device_noise_emulated(emu_lib):
    P.check_golden_device_noise(emu_lib)


@pytest.mark.gpu
def test_golden_device_noise_gpu(cuda_lib):
    P.check_golden_device_noise(cuda_lib)
