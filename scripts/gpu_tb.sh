#!/bin/bash
# temporal blocking: GPU parity + timings of the TBX_T variants (variants/libeq_T*.so)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tb_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tb_tests.log
tail -3 gpurun_out/tb_tests.log
for T in 1 2 3 4; do
  for O in 0 2; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T /"
  done
  EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 4096 20 1 3 2>&1 | tail -1 | sed "s/^/T=$T /"
done | tee gpurun_out/tb_times.log
