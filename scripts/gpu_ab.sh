#!/bin/bash
mkdir -p gpurun_out
for V in prev new prev new; do
  EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/$V /"
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/$V nodeps /"
done | tee gpurun_out/ab_times.log
for V in prev new; do
  EQ_LSX_CTAS_PER_SM=4 EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/$V ctas=4 /"
  EQ_LSX_CTAS_PER_SM=4 EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/$V ctas=4 nodeps /"
  EQ_LSX_CTAS_PER_SM=1 EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1 | sed "s/^/$V ctas=1 nodeps /"
done | tee -a gpurun_out/ab_times.log
