#!/bin/bash
mkdir -p gpurun_out
for PB in 1 2 4; do
for O in 0 1 2; do
  EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/tb pub=$PB /"
done
EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | sed "s/^/tb pub=$PB /"
done | tee gpurun_out/tb6_times.log
EQ_LSX_PUBBATCH=2 EQ_LSX_TRACE=0,0,0 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_tb_g0q0.log 2>&1
EQ_LSX_PUBBATCH=2 EQ_LSX_JOBTIMES=gpurun_out/jt_lean.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
