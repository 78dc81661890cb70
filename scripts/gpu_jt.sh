#!/bin/bash
mkdir -p gpurun_out
EQ_LSX_JOBTIMES=gpurun_out/jt_passive.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
EQ_LSX_CTAS_PER_SM=3 EQ_LSX_JOBTIMES=gpurun_out/jt_passive_c3.bin timeout 300 python scripts/prof_linsolve.py 16384 20 2 1 2>&1 | tail -1
