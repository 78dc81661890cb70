#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for args in "4096 1 2" "4096 40 2" "4096 40 0" "4096 40 1" "16384 1 2" "16384 20 2" "16384 20 1" "1024 20 2"; do
  timeout 120 python scripts/prof_linsolve.py $args 3
done 2>&1 | tee gpurun_out/linsolve_times_v4.txt
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c4_v4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_exact -s 1 -c 1 \
   -o gpurun_out/prof_lsx_r1d -f python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/ncu_lsx4.log 2>&1
tail -2 gpurun_out/ncu_lsx4.log
