#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['phases_ms_per_step'])" | tee gpurun_out/tb8.log
for NK in "8192 20" "4096 40" "2048 20"; do
 for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py $NK $O 3 2>&1 | tail -1 | sed "s/^/tb /"
  EQ_LSX_TB=0 timeout 300 python scripts/prof_linsolve.py $NK $O 3 2>&1 | tail -1 | sed "s/^/v7 /"
 done
done | tee -a gpurun_out/tb8.log
