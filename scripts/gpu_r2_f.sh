#!/bin/bash
mkdir -p gpurun_out
export EQ_RB_KERNEL=slide
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_slide -s 2 -c 1 -o gpurun_out/r2f_rbs python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
