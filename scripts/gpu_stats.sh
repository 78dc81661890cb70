#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for args in "4096 1 2" "4096 40 2" "4096 40 1" "16384 1 2" "16384 20 2" "16384 20 1" "16384 20 0" "1024 20 2"; do
  timeout 120 python scripts/prof_linsolve.py $args 3
done 2>&1 | tee gpurun_out/linsolve_times_latest.txt
EQ_LSX_TRACE=1 timeout 120 python scripts/prof_linsolve.py 16384 20 1 1 2>&1 | grep "b=0\|b=1\|N=" | cut -c1-200 | tee gpurun_out/trace_latest.txt
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c4_latest.json
