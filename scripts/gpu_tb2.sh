#!/bin/bash
mkdir -p gpurun_out
for T in 2 3 4; do
  for O in 0 1 2; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T /"
  done
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=$T nodeps /"
  EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 4096 20 1 3 2>&1 | tail -1 | sed "s/^/T=$T /"
done | tee gpurun_out/tb2_times.log
EQUILIBRIUM_CUDA_LIB=variants/libeq_T2.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_tb -s 1 -c 1 \
   -o gpurun_out/prof_tb_b -f python scripts/prof_linsolve.py 8192 20 2 1 > gpurun_out/ncu_tb.log 2>&1
tail -2 gpurun_out/ncu_tb.log
