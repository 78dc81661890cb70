#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for args in "16384 20 2 3 red_black" "16384 20 1 3 red_black" "4096 40 2 3 red_black" "1024 20 2 3 red_black"; do timeout 120 python scripts/prof_linsolve.py $args; done 2>&1 | tee gpurun_out/rb_times.txt
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c4_latest.json | cut -c1-200
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c3_latest.json | cut -c1-200
