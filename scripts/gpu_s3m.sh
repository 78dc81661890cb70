#!/bin/bash
# session 3, run M: which change made the red-black frame slower (62.7 -> 71.6 ms)?
mkdir -p gpurun_out
for V in E G NOPUSH; do
  EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 200 python scripts/prof_frame.py c4 red_black 3 2>&1 | tail -1 | sed "s/^/$V /"
  EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 200 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/$V /"
done | tee gpurun_out/s3m.log
timeout 200 python scripts/prof_frame.py c4 red_black 3 2>&1 | tail -1 | sed "s/^/HEAD /" | tee -a gpurun_out/s3m.log
timeout 200 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/HEAD /" | tee -a gpurun_out/s3m.log
