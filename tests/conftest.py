import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CUDA_LIB = os.path.join(ROOT, "equilibrium_b200", "libequilibrium_cuda.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libequilibrium_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (tests are one of the three places allowed to load it)."""
    from oracle import loader
    loader.build()
    return loader


@pytest.fixture(scope="session")
def emu_lib():
    """Product sources compiled against the host SIMT emulator (tests/emu).  EQUILIBRIUM_EMU_LIB selects another
    emulated build of the same sources (kernel variants under test, e.g. -DRBR_INLINE_BARRIER=1)."""
    os.environ.setdefault("EQ_EMU_SMS", "4")
    if os.environ.get("EQUILIBRIUM_EMU_LIB"):
        return os.path.abspath(os.environ["EQUILIBRIUM_EMU_LIB"])
    csrc = os.path.join(ROOT, "equilibrium_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    srcs += [os.path.join(ROOT, "tests", "emu", f) for f in ("cuda_emu.h", "cuda_emu.cpp")]
    srcs.append(os.path.join(ROOT, "include", "equilibrium_cuda.h"))
    if _newer(EMU_LIB, srcs):
        subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    return EMU_LIB


@pytest.fixture(scope="session")
def cuda_lib():
    """The real library; GPU tests must run through it (no fallback).  EQUILIBRIUM_CUDA_LIB selects another build of
    the same sources (kernel variants under test)."""
    path = os.path.abspath(os.environ.get("EQUILIBRIUM_CUDA_LIB") or CUDA_LIB)
    assert os.path.exists(path), "build libequilibrium_cuda.so first (__graft_entry__.build())"
    from equilibrium_b200 import _lib
    lib = _lib.load(path)
    assert lib.eq_device_count() > 0, "no CUDA device visible"
    return path
