#!/bin/bash
# session 3, run K (8 GPUs): config-5 sweep with a non-solenoidal initial state
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/c5_sweep.py --out gpurun_out/s3k_c5.json 2> gpurun_out/s3k_c5.err | tee gpurun_out/s3k.log
tail -3 gpurun_out/s3k_c5.err
