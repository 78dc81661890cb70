#!/bin/bash
mkdir -p gpurun_out
EQ_LSX_TRACE=1 timeout 120 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | cut -c1-200 | tee gpurun_out/trace_passive_k20.txt
