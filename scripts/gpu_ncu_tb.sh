#!/bin/bash
mkdir -p gpurun_out
EQUILIBRIUM_CUDA_LIB=variants/libeq_T2.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_tb -s 1 -c 1 \
   -o gpurun_out/prof_tb_a -f python scripts/prof_linsolve.py 8192 20 2 1 > gpurun_out/ncu_tb.log 2>&1
tail -2 gpurun_out/ncu_tb.log
