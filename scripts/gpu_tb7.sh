#!/bin/bash
mkdir -p gpurun_out
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1
done | tee gpurun_out/tb7_times.log
EQ_LSX_NODEPS=1 timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/nodeps /" | tee -a gpurun_out/tb7_times.log
for NK in "8192 20" "4096 40" "4096 20" "1024 20"; do
  timeout 300 python scripts/prof_linsolve.py $NK 2 3 2>&1 | tail -1 | sed "s/^/tb /"
  EQ_LSX_TB=0 timeout 300 python scripts/prof_linsolve.py $NK 2 3 2>&1 | tail -1 | sed "s/^/v7 /"
done | tee -a gpurun_out/tb7_times.log
EQ_LSX_JOBTIMES=gpurun_out/jt_final.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
