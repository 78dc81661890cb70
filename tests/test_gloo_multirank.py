"""CPU-only, world_size 2 over gloo: one PROCESS per rank, exactly how bench.py is launched on
GPUs (torch.distributed is only the plumbing that carries the rendezvous blobs).  The ranks run
the emulated build of the product sources and map each other's "device" memory through the
emulator's cudaIpc stand-ins (POSIX shared memory); the assembled slabs must be bit-identical
to the single-domain oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import parity as P
from conftest import ROOT

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["EQ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["EQ_ROOT"], "tests"))
import torch.distributed as dist
import parity as P
from equilibrium_b200 import Fluid, FluidConfigs, Rectangle, SimulationConfigs, connect_distributed

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, k, frames = int(os.environ["EQ_N"]), int(os.environ["EQ_K"]), int(os.environ["EQ_FRAMES"])
rects = json.loads(os.environ["EQ_RECTS"])
f = Fluid(FluidConfigs(), SimulationConfigs(0.02, k, n), lib_path=os.environ["EQ_EMU_LIB"], rank=rank, world=world)
connect_distributed(f)
for r in rects:
    f.fill_obstacle(Rectangle((r[0], r[1]), (r[2], r[3]), n))
f.step_n(frames, P.impulses(n, frames, 11))
f.sync()
out = {}
for name, _ in P.F32_FIELDS:
    r0, rows = f.download_owned(name)
    out[name] = rows
    out["row0"] = np.array([r0])
np.savez(os.path.join(os.environ["EQ_OUT"], f"rank{rank}.npz"), **out)
dist.barrier()
f.close()
dist.destroy_process_group()
'''


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_process_slabs_match_oracle(oracle, emu_lib, tmp_path):
    n, k, frames = 96, 2, 2
    rects = [[20, 28, 40, 40], [50, 60, 70, 66]]
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, EQ_ROOT=ROOT, EQ_EMU_LIB=emu_lib, EQ_N=str(n), EQ_K=str(k), EQ_FRAMES=str(frames),
               EQ_RECTS=str(rects), EQ_OUT=str(tmp_path), EQ_EMU_SMS="2", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(worker)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    ref = oracle.RefFluid(n, 0.02, k)
    for r in rects:
        ref.fill_rect(*r)
    for (_, x, y, ax, ay) in P.impulses(n, frames, 11):
        ref.add_velocity(x, y, ax, ay)
        ref.step()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert int(parts[0]["row0"][0]) == 0
    for name, fid in P.F32_FIELDS:
        got = np.concatenate([parts[0][name], parts[1][name]], axis=0)
        want = ref.field(fid)
        assert got.shape == want.shape
        assert P.bits_equal(got, want), f"{name}: {P.describe_diff(got, want)}"
