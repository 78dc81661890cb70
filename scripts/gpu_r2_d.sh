#!/bin/bash
mkdir -p gpurun_out
EQ_WF_DEBUG=1 timeout 120 python scripts/prof_linsolve.py 4096 16 2 1 > gpurun_out/r2d_dbg.log 2>&1
tail -2 gpurun_out/r2d_dbg.log
