"""Error behaviour of the C ABI (SURVEY 8b "Errors": every export returns 0 or a negative code, the message comes
from eq_last_error, nothing unwinds across the boundary).  Runs on the emulated build of the product sources: the
validation code is host code and is the same in both builds."""
import ctypes as C

import numpy as np
import pytest

import parity as P
from equilibrium_b200 import EquilibriumError, Fluid, FluidConfigs, SimulationConfigs, _lib


def mk(emu_lib, n=64, k=2, **kw):
    return Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=emu_lib, **kw)


@pytest.mark.parametrize("n", [0, 19, 32769])
def test_size_out_of_range(emu_lib, n):
    # init_density computes size/2 - 10 in u32 (fluid.rs:534): the reference itself cannot build a Fluid below 20
    with pytest.raises(EquilibriumError) as e:
        mk(emu_lib, n=n)
    assert "size" in str(e.value) and e.value.code < 0


def test_smallest_size_steps_like_the_oracle(oracle, emu_lib):
    dev, ref = P.make_pair(oracle, emu_lib, 20, 2, [(5, 5, 9, 8)])
    for _ in range(2):
        dev.step()
        ref.step()
    P.assert_state_equal(dev, ref, "N=20")


def test_bad_creation_parameters(emu_lib):
    lib = _lib.load(emu_lib)
    h = C.c_void_p()
    assert lib.eq_create(None, C.byref(h)) < 0 and b"null" in lib.eq_last_error()
    p = _lib.EqParams()
    p.size, p.delta_t, p.frames, p.mode = 64, 0.02, 4, 7
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0 and b"mode" in lib.eq_last_error()
    p.mode, p.frames = 0, -1
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0 and b"negative" in lib.eq_last_error()
    p.frames, p.world, p.rank = 4, 9, 0
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0
    p.world, p.rank = 2, 2
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0 and b"rank" in lib.eq_last_error()
    p.world, p.rank, p.size = 4, 0, 64        # 62 interior rows = 2 bands of 32 < 4 slabs
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0 and b"slab" in lib.eq_last_error()
    p.world, p.device = 1, 99
    assert lib.eq_create(C.byref(p), C.byref(h)) < 0 and b"device" in lib.eq_last_error()


def test_null_handle_is_an_error_not_a_crash(emu_lib):
    lib = _lib.load(emu_lib)
    for call in (lambda: lib.eq_step(None), lambda: lib.eq_sync(None), lambda: lib.eq_init_default(None),
                 lambda: lib.eq_step_n(None, 1, None, 0), lambda: lib.eq_step_n_noise(None, 1, None),
                 lambda: lib.eq_fill_rect(None, 1, 1, 2, 2), lambda: lib.eq_op_add_source(None, 0, 1, 1.0)):
        assert call() < 0
        assert lib.eq_last_error()


def test_field_transfers_check_their_sizes(emu_lib):
    f = mk(emu_lib)
    lib = f._lib
    a = np.zeros((64, 64), dtype=np.float32)
    assert lib.eq_upload(f._h, _lib.F_DENSITY, a.ctypes.data, a.nbytes - 4) < 0
    assert lib.eq_download(f._h, _lib.F_DENSITY, a.ctypes.data, a.nbytes + 4) < 0
    assert lib.eq_download(f._h, _lib.F_CELLS, a.ctypes.data, a.nbytes) < 0          # cells are 1 byte each
    assert lib.eq_download(f._h, 17, a.ctypes.data, a.nbytes) < 0
    assert lib.eq_upload(f._h, _lib.F_DENSITY, None, a.nbytes) < 0
    assert lib.eq_download_rows(f._h, _lib.F_DENSITY, 60, 8, a.ctypes.data) < 0       # rows 60..67 leave the grid


def test_operator_argument_checks(emu_lib):
    f = mk(emu_lib)
    with pytest.raises(EquilibriumError):
        f.op_lin_solve(P.ROW, "velocities_x", "velocities_x", 0.1, 1.4, 2)           # x aliases x0
    with pytest.raises(EquilibriumError):
        f.op_lin_solve(5, "velocities_x", "velocities_x0", 0.1, 1.4, 2)              # no such orientation
    with pytest.raises(EquilibriumError):
        f.op_lin_solve(P.ROW, "velocities_x", "velocities_x0", 0.1, 1.4, -1)
    with pytest.raises(EquilibriumError):
        f.op_lin_solve(P.ROW, "cells_type", "velocities_x0", 0.1, 1.4, 1)            # not an f32 field
    with pytest.raises(EquilibriumError):
        f.op_set_boundaries(3, "density")
    lib = f._lib
    assert lib.eq_step_n(f._h, -1, None, 0) < 0 and lib.eq_step_n(f._h, 1, None, 2) < 0
    assert lib.eq_step_n_noise(f._h, 1, None) < 0 and lib.eq_step_n_noise(f._h, -1, C.byref(_lib.EqNoise())) < 0


def test_set_params_cannot_resize(emu_lib):
    # the reference builds a new Fluid for a new size (renderer.rs:145-149)
    f = mk(emu_lib)
    f.simulation_configs.size = 128
    with pytest.raises(EquilibriumError):
        f.step()
    f.simulation_configs.size = 64
    f.simulation_configs.delta_t = 0.05            # everything else is a live setter
    f.fluid_configs.viscousity = 0.01
    f.step()
    f.sync()


def test_snapshot_argument_checks(emu_lib):
    f = mk(emu_lib)
    out = np.zeros((64, 64), dtype=np.float32)
    with pytest.raises(EquilibriumError):
        f.snapshot_begin(out, slot=_lib.SNAPSHOT_SLOTS)
    with pytest.raises(EquilibriumError):
        f.snapshot_begin(out[:32], slot=0)                                          # half a frame
    with pytest.raises(EquilibriumError):
        f.snapshot_wait(-1)
    f.snapshot_begin(out, slot=1)
    f.snapshot_wait(1)
    assert P.bits_equal(out, f.download("density"))


def test_errors_do_not_poison_the_handle(oracle, emu_lib):
    dev, ref = P.make_pair(oracle, emu_lib, 64, 2)
    with pytest.raises(EquilibriumError):
        dev.op_add_source("density", "density", 1.0)
    dev.step()
    ref.step()
    P.assert_state_equal(dev, ref, "after a rejected call")
