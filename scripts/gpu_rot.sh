#!/bin/bash
# role rotation A/B on v7 (T=1) and TB (T=2); nodeps for TB
mkdir -p gpurun_out
for T in 1 2; do
 for ROT in 0 1; do
  for O in 0 2; do
    EQ_LSX_ROT=$ROT EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T rot=$ROT /"
  done
 done
done | tee gpurun_out/rot_times.log
for T in 1 2; do
 for C in 4 3 2; do
  EQ_LSX_CTAS_PER_SM=$C EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 0 3 2>&1 | tail -1 | sed "s/^/T=$T rot=1 ctas=$C /"
 done
 EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 0 3 2>&1 | tail -1 | sed "s/^/T=$T rot=1 nodeps /"
 EQ_LSX_ROT=0 EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 0 3 2>&1 | tail -1 | sed "s/^/T=$T rot=0 nodeps /"
done | tee -a gpurun_out/rot_times.log
