#!/bin/bash
# ncu of k_linsolve_wf without dependency waits, one job per SM: the compute warp's own speed
mkdir -p gpurun_out
export EQ_EXACT_KERNEL=wf EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg.so EQ_LSX_NODEPS=1 EQ_WF_CTAS_PER_SM=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_wf -s 1 -c 1 -o gpurun_out/r2f_wf1 python scripts/prof_linsolve.py 4096 4 2 1 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
