#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rb_stream -s 2 -c 1 -o gpurun_out/r02_rb_stream_c4 -f python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/r2q_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rb_stream -s 2 -c 1 -o gpurun_out/r02_rb_stream_c4_row -f python scripts/prof_linsolve.py 16384 20 0 1 red_black > gpurun_out/r2q_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/r2q_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu1.log; wc -l gpurun_out/r02_launches_c4.csv
