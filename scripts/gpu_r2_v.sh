#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "16384 20 0" "16384 20 1" "16384 20 2" "4096 40 0" "4096 40 1" "4096 40 2"; do
  set -- $cfg
  timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2v.log
scripts/gpu_r2_final.sh
