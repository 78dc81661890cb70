#!/bin/bash
mkdir -p gpurun_out
EQ_LSX_TRACE=0,100,0 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_tb_mid_q0.log 2>&1
EQ_LSX_TRACE=0,100,990 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_tb_mid_qend.log 2>&1
EQ_LSX_JOBTIMES=gpurun_out/jt_final.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
