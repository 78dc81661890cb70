#!/bin/bash
# session 3, run O (8 GPUs): halo push in its own branch -- parity, bench at N=8, config-5 sweep (viscosity 0)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/s3o.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/s3o_bench_c4_n8.json 2> gpurun_out/s3o_bench_c4_n8.err
python - <<'PY' | tee -a gpurun_out/s3o.log
import json
try:
    d = json.loads(open("gpurun_out/s3o_bench_c4_n8.json").read().strip().splitlines()[-1])
    rb = d.get("red_black") or {}
    print("n8 exact ms/step", d["ms_per_step"], "rb ms/step", rb.get("ms_per_step"), "rb phases", rb.get("phases_ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("n8 failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/c5_sweep.py --out gpurun_out/s3o_c5.json 2> gpurun_out/s3o_c5.err | tee -a gpurun_out/s3o.log
