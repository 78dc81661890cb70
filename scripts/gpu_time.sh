#!/bin/bash
mkdir -p gpurun_out
for args in "4096 1 2" "4096 40 2" "4096 40 0" "4096 40 1" "16384 1 2" "16384 20 2" "16384 20 1" "1024 20 2"; do
  timeout 120 python scripts/prof_linsolve.py $args 3
done 2>&1 | tee gpurun_out/linsolve_times_latest.txt
