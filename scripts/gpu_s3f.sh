#!/bin/bash
# session 3, run F (2 GPUs): row-slab parity tests with the new kernels, bench at N=2, S16 ring variant on one GPU
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_red_black.py -x -q 2>&1 | tail -5 ) | tee gpurun_out/s3f.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s3f_bench_c4_n2.json 2> gpurun_out/s3f_bench_c4_n2.err
tail -3 gpurun_out/s3f_bench_c4_n2.err
python - <<'PY' | tee -a gpurun_out/s3f.log
import json
try:
    d = json.loads(open("gpurun_out/s3f_bench_c4_n2.json").read().strip().splitlines()[-1])
    print("n2 ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"))
except Exception as e:
    print("n2 failed", e)
PY
for O in 0 2; do
EQUILIBRIUM_CUDA_LIB=variants/libeq_S16.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/S16 /" | tee -a gpurun_out/s3f.log
done
EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_S16.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/S16 nodeps /" | tee -a gpurun_out/s3f.log
EQUILIBRIUM_CUDA_LIB=variants/libeq_S16.so timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 2>&1 | tail -1 | sed "s/^/S16 /" | tee -a gpurun_out/s3f.log
