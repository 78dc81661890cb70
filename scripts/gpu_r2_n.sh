#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_red_black.py -x -q -m gpu 2>&1 | tail -3 ) | tee gpurun_out/r2n_pytest.log
{
for cfg in "16384 20 2" "16384 20 0" "4096 40 2"; do
  set -- $cfg
  timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1 | sed "s/^/default $cfg: /"
done
} 2>&1 | tee gpurun_out/r2n.log
timeout 900 python bench.py --no-configs > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 3000 gpurun_out/r2n_bench.json
