#!/usr/bin/env python
"""BASELINE config 5: 32768^2 on 8 GPUs (row slabs), Gauss-Seidel iteration sweep K = 20 / 80 / 200 with the
pressure-solve convergence logged.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        scripts/c5_sweep.py [--size 32768] [--ks 20,80,200] [--mode red_black] [--frames 2] [--out gpurun_out/c5.json]

Per K: a fresh fluid in an analytic, NON-solenoidal initial state (vx = sin(2 pi x / N), vy = sin(2 pi y / N): its
divergence is (2 pi / N)(cos + cos), so the pressure solve has real work to do -- the reference's uniform v = (1,1)
start and SURVEY 8d's sin(y), sin(x) pair are divergence-free away from the walls) with a checker-block density,
`frames` frames timed on the device (max over ranks), then
  * div_l2      = || div(u) ||_2 of the velocity the step leaves behind (eq_divergence_l2, stencil of fluid.rs:341-345)
  * gs_residual = || 4 p - (sum of the 4 neighbours) - div ||_2 of the LAST pressure solve (project #2 keeps p in
                  velocities_x0 and div in velocities_y0, fluid.rs:491-499), over the rows each rank owns minus the two
                  slab-edge rows, computed on the host from the downloaded rows
both reduced over the ranks.  Rank 0 prints one JSON line per K and writes the list to --out.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--ks", default="20,80,200")
    ap.add_argument("--mode", default="red_black", choices=["red_black", "exact"])
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--viscosity", type=float, default=0.0,
                    help="0 (default): diffusion is the identity (a = 0) and the velocity survives, so the projection has "
                         "work to do; with the reference's 0.001 the diffusion number dt*visc*(N-2)^2 is 21 474 at this size "
                         "and K sweeps from the stale zero guess (quirk Q3) wipe the velocity out in one frame")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c5.json"))
    args = ap.parse_args()
    from equilibrium_b200 import Fluid, FluidConfigs, SimulationConfigs, connect_distributed

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:                                     # torch.distributed is plumbing only: rendezvous blobs and reductions
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def reduce(vals, op):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]
    n = args.size
    results = []
    for k in [int(x) for x in args.ks.split(",")]:
        f = Fluid(FluidConfigs(diffusion=0.0, viscousity=args.viscosity), SimulationConfigs(0.02, k, n), mode=args.mode,
                  device=local, rank=rank, world=world)
        if world > 1:
            connect_distributed(f)
        r0, r1 = f.owned_rows()
        rows = r1 - r0
        # analytic initial state on the owned rows
        x = np.arange(n, dtype=np.float64)
        y = np.arange(r0, r1, dtype=np.float64)
        vx = np.broadcast_to(np.sin(2 * np.pi * x / n)[None, :], (rows, n)).astype(np.float32)
        vy = np.broadcast_to(np.sin(2 * np.pi * y / n)[:, None], (rows, n)).astype(np.float32)
        blk = max(1, n // 64)
        dens = ((((np.arange(r0, r1) // blk)[:, None] + (np.arange(n) // blk)[None, :]) & 1) * 0.9).astype(np.float32)
        lib = f._lib
        from equilibrium_b200 import _lib as L
        for name, arr in (("velocities_x", vx), ("velocities_y", vy), ("density", dens), ("scratch_space", dens)):
            a = np.ascontiguousarray(arr)
            L.check(lib, lib.eq_upload_rows(f._h, f.FIELDS[name], r0, rows, a.ctypes.data))
        del vx, vy, dens
        div0 = reduce([f.divergence_l2() ** 2], "sum")[0] ** 0.5      # || div(u) ||_2 of the initial state
        f.step_n(1)                                   # warm-up frame (tables, first-touch)
        f.sync()
        if world > 1:
            dist.barrier()
        f.timer_start()
        f.step_n(args.frames)
        ms = f.timer_stop()
        f.sync()
        ms = reduce([ms], "max")[0]
        d2 = f.divergence_l2() ** 2
        # residual of the last pressure solve on the owned rows (minus the slab-edge rows, whose vertical
        # neighbours live on another rank)
        p = np.empty((rows, n), dtype=np.float32)
        dv = np.empty((rows, n), dtype=np.float32)
        L.check(lib, lib.eq_download_rows(f._h, f.FIELDS["velocities_x0"], r0, rows, p.ctypes.data))
        L.check(lib, lib.eq_download_rows(f._h, f.FIELDS["velocities_y0"], r0, rows, dv.ctypes.data))
        res2 = 0.0
        for j0 in range(1, rows - 1, 256):            # in strips: float64 temporaries of a 4 GiB field would not fit
            j1 = min(rows - 1, j0 + 256)
            pc = p[j0:j1, 1:-1].astype(np.float64)
            s = (p[j0:j1, 2:].astype(np.float64) + p[j0:j1, :-2] + p[j0 + 1:j1 + 1, 1:-1] + p[j0 - 1:j1 - 1, 1:-1])
            r = 4.0 * pc - s - dv[j0:j1, 1:-1]
            res2 += float(np.sum(r * r))
        del p, dv
        d2, res2 = reduce([d2, res2], "sum")
        rec = {"config": "c5", "size": n, "gs_iterations": k, "viscosity": args.viscosity, "mode": args.mode, "n_gpus": world, "frames": args.frames,
               "ms_per_frame": ms / args.frames, "cell_updates_per_s": n * n * args.frames / (ms * 1e-3),
               "cell_updates_per_s_per_gpu": n * n * args.frames / (ms * 1e-3) / world,
               "div_l2_initial": div0, "div_l2_after_step": d2 ** 0.5, "gs_residual_l2_last_pressure_solve": res2 ** 0.5}
        results.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        f.close()
        if world > 1:
            dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(results, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
