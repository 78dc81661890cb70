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


_cuda_probe = {}


def _probe_cuda():
    """(path, n_devices, reason) of the product library, probed once.  Never asserts: a session fixture that raises
    takes every GPU test with it and hides the CUDA error that caused it (round-1 driver record)."""
    if _cuda_probe:
        return _cuda_probe["v"]
    path = os.path.abspath(os.environ.get("EQUILIBRIUM_CUDA_LIB") or CUDA_LIB)
    n, why = 0, ""
    if not os.path.exists(path):
        why = f"{path} is missing: build it first (__graft_entry__.build())"
    else:
        from equilibrium_b200 import _lib
        lib = _lib.load(path)
        n = lib.eq_device_count()          # retries transient start-up errors itself (EQ_DEVICE_WAIT_S)
        if n <= 0:
            why = (lib.eq_last_error() or b"").decode("utf-8", "replace") or "eq_device_count() == 0"
            try:
                smi = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
                why += f" | nvidia-smi -L rc={smi.returncode}: {(smi.stdout + smi.stderr).strip()[:400]}"
            except Exception as e:   # noqa: BLE001 - diagnostics only
                why += f" | nvidia-smi not runnable: {e}"
            why += f" | CUDA_VISIBLE_DEVICES={os.environ.get('CUDA_VISIBLE_DEVICES', '<unset>')}"
    _cuda_probe["v"] = (path, n, why)
    return _cuda_probe["v"]


@pytest.fixture(scope="session")
def cuda_lib():
    """The real library; GPU tests must run through it (no fallback).  EQUILIBRIUM_CUDA_LIB selects another build of
    the same sources (kernel variants under test)."""
    path, n, why = _probe_cuda()
    if n <= 0:
        pytest.fail(f"GPU test needs a CUDA device and there is no CPU fallback: {why}", pytrace=False)
    return path


@pytest.fixture(scope="session")
def cuda_devices(cuda_lib):
    """Number of visible CUDA devices (multi-GPU tests skip themselves when it is too small)."""
    return _probe_cuda()[1]


# Order of the test files under `-x`: single-GPU kernel parity first (the record that matters most), the multi-GPU
# row-slab tests after it, host-mirror frame loops last.  Files not listed keep their alphabetical place in between.
_FILE_ORDER = ["test_oracle.py", "test_abi.py", "test_gpu_parity.py", "test_red_black.py", "test_sources.py"]
_FILE_LAST = ["test_gpu_multigpu.py", "test_zz_host_mirrors.py"]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.basename(str(item.fspath))
        if name in _FILE_ORDER:
            return (0, _FILE_ORDER.index(name))
        if name in _FILE_LAST:
            return (2, _FILE_LAST.index(name))
        return (1, 0)
    items.sort(key=key)   # stable: keeps the order inside a file
