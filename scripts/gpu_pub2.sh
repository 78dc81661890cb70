#!/bin/bash
mkdir -p gpurun_out
for PB in 6 8 12 16; do
  EQ_LSX_TB=0 EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/v7 pub=$PB /"
  EQ_LSX_TB=0 EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | sed "s/^/v7 pub=$PB /"
done | tee gpurun_out/pub2_times.log
for PB in 2 4; do
  EQUILIBRIUM_CUDA_LIB=variants/libeq_T3.so EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T3 pub=$PB /"
done | tee -a gpurun_out/pub2_times.log
for O in 0 1; do
EQ_LSX_TB=0 EQ_LSX_PUBBATCH=4 timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/v7 pub=4 /"
EQ_LSX_PUBBATCH=2 timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/tb pub=2 /"
done | tee -a gpurun_out/pub2_times.log
EQ_LSX_PUBBATCH=2 EQ_LSX_JOBTIMES=gpurun_out/jt_pub2.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
EQ_LSX_TB=0 EQ_LSX_PUBBATCH=4 EQ_LSX_TRACE=1 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_v7_pub4.log 2>&1
