#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tb_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tb_tests.log
tail -3 gpurun_out/tb_tests.log
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1
done | tee gpurun_out/tb3_times.log
timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | tee -a gpurun_out/tb3_times.log
EQ_LSX_TB=0 timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | sed "s/^/v7 /" | tee -a gpurun_out/tb3_times.log
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/tb3_bench.log
