#!/bin/bash
mkdir -p gpurun_out
EQ_LSX_TRACE=1 timeout 120 python scripts/prof_linsolve.py 16384 1 2 1 2>&1 | tail -34 | tee gpurun_out/trace_k1.txt
EQ_LSX_TRACE=1 timeout 120 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -34 | tee gpurun_out/trace_k20.txt
