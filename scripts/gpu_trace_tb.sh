#!/bin/bash
mkdir -p gpurun_out
EQ_LSX_PUBBATCH=2 EQ_LSX_TRACE=0,0,0 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_tb_g0q0.log 2>&1
EQ_LSX_PUBBATCH=2 EQ_LSX_TRACE=0,100,0 timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/trace_tb_g0b100q0.log 2>&1
tail -1 gpurun_out/trace_tb_g0b100q0.log
