#!/bin/bash
# launch list of the bench command on the final sources
mkdir -p gpurun_out
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-configs --no-extras > gpurun_out/r2w_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_c4.csv
