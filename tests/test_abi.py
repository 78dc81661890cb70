"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares,
the Python mirror declares a prototype for each, and host-side validation behaves
like the reference.  No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from equilibrium_b200 import Rectangle, SimulationConfigs, FluidConfigs, _lib
from conftest import CUDA_LIB, ROOT

HEADER = os.path.join(ROOT, "include", "equilibrium_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eq_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from equilibrium_b200 import build
    build.build()            # no-op unless the sources are newer than the library
    lib = C.CDLL(CUDA_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {CUDA_LIB}"
    assert _lib.load(CUDA_LIB).eq_abi_version() == 1


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.EqParams) == 4 + 4 + 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 128
    assert C.sizeof(_lib.EqSource) == 8 + 4 + 4 + 4 + 4 + 4 + 4   # padded to 8
    assert C.sizeof(_lib.EqProfile) == 12 * 8


def test_rect_valid_needs_no_device():
    lib = _lib.load(CUDA_LIB)
    assert lib.eq_rect_valid(80, 80, 110, 110, 128) == 1
    assert lib.eq_rect_valid(50, 120, 127, 110, 128) == 0      # obstacle.rs:100-107
    assert lib.eq_rect_valid(12, 12, 10, 10, 128) == 0         # obstacle.rs:109-115


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        _lib.load(str(tmp_path / "nope.so"))


def test_rectangle_panics_like_the_reference():
    with pytest.raises(ValueError):
        Rectangle((50, 120), (127, 110), 128)
    with pytest.raises(ValueError):
        Rectangle((12, 12), (10, 10), 128)
    r = Rectangle.default()
    assert r.get_approximate_points() == [(80, 80), (110, 110)]


def test_config_defaults_match_reference():
    s, f = SimulationConfigs(), FluidConfigs()
    assert (s.delta_t, s.frames, s.size) == (0.02, 16, 128)            # configs.rs:14-22
    assert (f.diffusion, f.viscousity, f.has_perlin_noise) == (0.0, 0.001, True)   # configs.rs:50-60


def test_rust_sys_crate_binds_only_declared_symbols():
    """rust/ cannot be compiled in this image (no cargo): at least keep its extern block inside the header."""
    rs = open(os.path.join(ROOT, "rust", "equilibrium-cuda-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (eq_[a-z0-9_]+)\s*\(", rs))
    assert bound and bound <= set(declared_symbols()), sorted(bound - set(declared_symbols()))
    for needed in ("eq_create", "eq_step", "eq_fill_rect", "eq_add_velocity", "eq_clone", "eq_destroy", "eq_download",
                   "eq_snapshot_begin", "eq_snapshot_wait"):
        assert needed in bound, needed


def test_colors_layout_matches_header():
    assert C.sizeof(_lib.EqColors) == 12 and _lib.SNAPSHOT_SLOTS == 2


def test_no_fused_multiply_add_outside_the_division_routine():
    """Bit-exactness needs every multiply and add rounded separately (rustc never contracts).  Static
    check of the compiled SASS: FFMA may only appear in the kernels that divide (the correctly rounded __fdiv_rn
    sequence of the divergence stencil, fluid.rs:341-345) -- never in a solver, advect or gradient kernel."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump") or not os.path.exists(CUDA_LIB):
        pytest.skip("needs cuobjdump and the built library")
    sass = subprocess.run(["cuobjdump", "-sass", CUDA_LIB], capture_output=True, text=True, timeout=300).stdout
    fn, with_fma, seen = None, set(), set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            seen.add(fn)
        elif fn and re.search(r"\bFFMA\b", line):
            with_fma.add(fn)
    assert any("k_linsolve_tb" in f for f in seen) and any("k_rb_reg" in f for f in seen) and any("k_advect" in f for f in seen)
    assert all("k_divergence" in f for f in with_fma), sorted(with_fma)
