#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/HEAD2 /" | tee gpurun_out/s3n.log
timeout 200 python scripts/prof_linsolve.py 16384 20 0 3 red_black 2>&1 | tail -1 | sed "s/^/HEAD2 /" | tee -a gpurun_out/s3n.log
timeout 200 python scripts/prof_frame.py c4 red_black 3 2>&1 | tail -1 | sed "s/^/HEAD2 /" | tee -a gpurun_out/s3n.log
