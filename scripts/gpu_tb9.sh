#!/bin/bash
mkdir -p gpurun_out
for T in 2 3 4; do
  for O in 0 1 2; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T /"
  done
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=$T nodeps /"
  EQ_LSX_PUBBATCH=2 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=$T pub=2 /"
done | tee gpurun_out/tb9.log
EQ_LSX_CTAS_PER_SM=4 EQUILIBRIUM_CUDA_LIB=variants/libeq_T2.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=2 ctas=4 /" | tee -a gpurun_out/tb9.log
EQUILIBRIUM_CUDA_LIB=variants/libeq_T4.so timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 2>&1 | tail -1 | sed "s/^/T=4 /" | tee -a gpurun_out/tb9.log
