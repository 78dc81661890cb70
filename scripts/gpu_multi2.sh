#!/bin/bash
# 2 GPUs: parity tests of the row-slab path, then bench c4 at N=2 (exact) with the release batch swept
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
for PB in 8 16; do
EQ_LSX_PUBBATCH=$PB timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pub=$PB', d['ms_per_step'], d['value'], d['roofline']['phases_ms_per_step'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
tail -c 1500 gpurun_out/bench_c4_n2.json
