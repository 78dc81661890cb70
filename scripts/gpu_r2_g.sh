#!/bin/bash
mkdir -p gpurun_out
export EQ_EXACT_KERNEL=wf
{
for cfg in "16384 20 2" "16384 20 0" "16384 20 1" "4096 40 2" "4096 40 0" "4096 40 1"; do
  set -- $cfg
  timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 2>&1 | tail -1 | sed "s/^/wf $cfg: /"
done
EQ_WF_CTAS_PER_SM=3 timeout 120 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/wf 3ctas: /"
export EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg.so
for C in 4 3 1; do
  EQ_LSX_NODEPS=1 EQ_WF_CTAS_PER_SM=$C timeout 120 python scripts/prof_linsolve.py 16384 20 2 2 2>&1 | tail -1 | sed "s/^/wf nodeps ctas=$C: /"
done
EQ_LSX_NODEPS=1 timeout 120 python scripts/prof_linsolve.py 4096 40 2 2 2>&1 | tail -1 | sed "s/^/wf nodeps c3: /"
unset EQUILIBRIUM_CUDA_LIB
} 2>&1 | tee gpurun_out/r2g.log
( timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2g_pytest.log
