#!/bin/bash
mkdir -p gpurun_out
export EQ_EXACT_KERNEL=wf
for V in t2c5 t2c4 t4c3; do
  for cfg in "16384 20 2" "16384 20 0" "4096 40 2"; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so timeout 60 python scripts/prof_linsolve.py $cfg 3 2>&1 | tail -1 | sed "s/^/$V: /"
  done
done 2>&1 | tee gpurun_out/r2d_variants.log
