#!/bin/bash
mkdir -p gpurun_out
for sl in 0 8 24 64 160; do echo "slack=$sl"; EQ_LSX_SLACK=$sl timeout 120 python scripts/prof_linsolve.py 16384 20 2 2; EQ_LSX_SLACK=$sl timeout 120 python scripts/prof_linsolve.py 4096 40 2 2; done 2>&1 | tee gpurun_out/sweep.txt
