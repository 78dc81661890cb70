#!/bin/bash
mkdir -p gpurun_out
for T in 3 4; do
  for O in 0 2; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T /"
  done
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=$T nodeps /"
done | tee gpurun_out/tb5_times.log
for C in 4 3; do
EQ_LSX_CTAS_PER_SM=$C timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=2 ctas=$C /"
done | tee -a gpurun_out/tb5_times.log
EQ_LSX_ROT=0 timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=2 rot=0 /" | tee -a gpurun_out/tb5_times.log
