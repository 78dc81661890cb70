#!/bin/bash
# session 3, run G: k_rb_reg without the precomputed guard predicates; launch list of the bench command for profiles/
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/s3g.log
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 red_black 2>&1 | tail -1 | sed "s/^/reg /"
done | tee -a gpurun_out/s3g.log
for O in 0 2; do
timeout 300 python scripts/prof_linsolve.py 4096 40 $O 3 red_black 2>&1 | tail -1 | sed "s/^/reg /" | tee -a gpurun_out/s3g.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3g_launches_c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/s3g_ncu_bench_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3g_launches_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-extras > gpurun_out/s3g_ncu_bench_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_reg -s 3 -c 1 \
   -o gpurun_out/s3g_rbreg_c4 -f python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/s3g_ncu_rb_c4.log 2>&1
ls -la gpurun_out/s3g* | tee -a gpurun_out/s3g.log
