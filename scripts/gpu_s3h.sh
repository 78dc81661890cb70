#!/bin/bash
# session 3, run H (8 GPUs): row-slab parity on 4 and 8 GPUs, bench at N=8 (exact + red-black + e2e)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -5 ) | tee gpurun_out/s3h.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/s3h_bench_c4_n8.json 2> gpurun_out/s3h_bench_c4_n8.err
tail -3 gpurun_out/s3h_bench_c4_n8.err
python - <<'PY' | tee -a gpurun_out/s3h.log
import json
try:
    d = json.loads(open("gpurun_out/s3h_bench_c4_n8.json").read().strip().splitlines()[-1])
    print("n8 ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"))
except Exception as e:
    print("n8 failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/s3h_bench_c4_n4.json 2> gpurun_out/s3h_bench_c4_n4.err
tail -3 gpurun_out/s3h_bench_c4_n4.err
python - <<'PY' | tee -a gpurun_out/s3h.log
import json
try:
    d = json.loads(open("gpurun_out/s3h_bench_c4_n4.json").read().strip().splitlines()[-1])
    print("n4 ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"))
except Exception as e:
    print("n4 failed", e)
PY
