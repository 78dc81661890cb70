#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_exact -s 1 -c 1 \
   -o gpurun_out/prof_lsx_r1e -f python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/ncu_lsx5.log 2>&1
tail -2 gpurun_out/ncu_lsx5.log
