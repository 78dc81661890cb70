#!/bin/bash
mkdir -p gpurun_out
echo "== 1 GPU kgroup sweep (lin_solve 16384 K=20 passive)"
for g in 20 5 2; do echo "kgroup=$g"; EQ_LSX_KGROUP=$g timeout 120 python scripts/prof_linsolve.py 16384 20 2 2; done
for g in 10 5 3 2 1; do
echo "== bench c4 N=2 kgroup=$g"
EQ_LSX_KGROUP=$g timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 2 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['phases_ms_per_step'])"
done
