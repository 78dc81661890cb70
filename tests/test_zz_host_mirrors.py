"""Host mirrors of the caller's side.  (Named to sort last: these drive whole frame loops, the kernel parity tests
run first.)

The C++ host mirror (include/equilibrium.hpp: the reference's `simulation` API above the C ABI) driven by
tests/cpp/host_mirror_test.cpp and bit-compared with the oracle: on the emulated build here, on the GPU under -m gpu."""
import os
import subprocess

import pytest

import parity as P
from conftest import ROOT


def build_and_run(tmp_path, lib_path, oracle_mod, env=None):
    oracle_lib = oracle_mod.lib()._name
    exe = str(tmp_path / "host_mirror_test")
    libdir, libname = os.path.split(os.path.abspath(lib_path))
    odir, oname = os.path.split(os.path.abspath(oracle_lib))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-o", exe, "-L", libdir, f"-l:{libname}", "-L", odir, f"-l:{oname}",
                           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "host mirror ok" in out.stdout


def test_cpp_host_mirror_on_the_emulated_build(tmp_path, oracle, emu_lib):
    build_and_run(tmp_path, emu_lib, oracle, env=dict(os.environ, EQ_EMU_SMS="4"))


@pytest.mark.gpu
def test_cpp_host_mirror_on_the_gpu(tmp_path, oracle, cuda_lib):
    build_and_run(tmp_path, cuda_lib, oracle)


@pytest.mark.gpu
def test_current_simulation_loop_on_the_gpu(oracle, cuda_lib):
    # the Python mirror of CurrentSimulation on the default scene's grid (the emulated run is in test_emu_parity.py)
    P.check_current_simulation(oracle, cuda_lib, n=128, k=6, rect=(80, 80, 110, 110))
