#!/bin/bash
mkdir -p gpurun_out
export EQ_EXACT_KERNEL=wf
{
for cfg in "16384 20 2" "16384 20 0" "16384 20 1" "4096 40 2" "4096 40 0" "4096 40 1" "4096 4 2" "1024 20 2"; do
  set -- $cfg
  timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 2>&1 | tail -1 | sed "s/^/wf $cfg: /"
done
EQ_WF_PUBBATCH=2 timeout 120 python scripts/prof_linsolve.py 16384 20 2 2 2>&1 | tail -1 | sed "s/^/wf pub2 16384 20 2: /"
EQ_WF_PUBBATCH=8 timeout 120 python scripts/prof_linsolve.py 16384 20 2 2 2>&1 | tail -1 | sed "s/^/wf pub8 16384 20 2: /"
EQ_WF_CTAS_PER_SM=2 timeout 120 python scripts/prof_linsolve.py 16384 20 2 2 2>&1 | tail -1 | sed "s/^/wf 2cta 16384 20 2: /"
} 2>&1 | tee gpurun_out/r2e.log
( timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2e_pytest.log
