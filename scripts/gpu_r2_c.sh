#!/bin/bash
# Round-2 call C (one GPU): first light of k_linsolve_wf -- parity suite, then lin_solve timings against k_linsolve_tb.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/r2c_pytest.log
{
for O in 0 1 2; do
  python scripts/prof_linsolve.py 16384 20 $O 3 | tail -1 | sed "s/^/wf: /"
  EQ_EXACT_KERNEL=tb python scripts/prof_linsolve.py 16384 20 $O 2 | tail -1 | sed "s/^/tb: /"
done
for O in 0 1 2; do
  python scripts/prof_linsolve.py 4096 40 $O 3 | tail -1 | sed "s/^/wf: /"
  EQ_EXACT_KERNEL=tb python scripts/prof_linsolve.py 4096 40 $O 2 | tail -1 | sed "s/^/tb: /"
done
EQ_WF_GENERAL=1 python scripts/prof_linsolve.py 4096 40 2 2 | tail -1 | sed "s/^/wf general only: /"
for C in 1 2; do EQ_WF_CTAS_PER_SM=$C python scripts/prof_linsolve.py 16384 20 2 2 | tail -1 | sed "s/^/wf ctas_per_sm=$C: /"; done
} 2>&1 | tee gpurun_out/r2c.log
