#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_red_black.py -x -q -m gpu 2>&1 | tail -3 ) | tee gpurun_out/r2n_pytest.log
{
for segs in 14 28 42 56 84; do
 for o in 0 1 2; do
  EQ_RQ_SEGS=$segs timeout 120 python scripts/prof_linsolve.py 16384 20 $o 3 red_black 2>&1 | tail -1 | sed "s/^/segs=$segs: /"
 done
done
for segs in 7 14 21 28; do
 for o in 0 2; do
  EQ_RQ_SEGS=$segs timeout 120 python scripts/prof_linsolve.py 4096 40 $o 3 red_black 2>&1 | tail -1 | sed "s/^/c3 segs=$segs: /"
 done
done
} 2>&1 | tee gpurun_out/r2n.log
timeout 900 python bench.py --no-configs > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 600 gpurun_out/r2n_bench.json
