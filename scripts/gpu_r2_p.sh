#!/bin/bash
# usage: gpu_r2_p.sh N   -- multi-GPU tests and the bench under torchrun on N GPUs
N=${1:-2}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2p_pytest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-configs > gpurun_out/r2p_bench_$N.json 2> gpurun_out/r2p_bench_$N.err
tail -c 300 gpurun_out/r2p_bench_$N.json
