#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "16384 20 2" "16384 20 0" "4096 40 2"; do
  set -- $cfg
  EQ_RB_KERNEL=slide timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1 | sed "s/^/slide $cfg: /"
  EQ_RB_KERNEL=slide EQUILIBRIUM_CUDA_LIB=variants/libeq_c5.so timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1 | sed "s/^/slide c5 $cfg: /"
  EQ_RB_KERNEL=slide EQUILIBRIUM_CUDA_LIB=variants/libeq_c6.so timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1 | sed "s/^/slide c6 $cfg: /"
done
} 2>&1 | tee gpurun_out/r2l.log
( EQ_RB_KERNEL=slide timeout 900 python -m pytest tests/test_red_black.py tests/test_gpu_parity.py -x -q -m gpu -k "red_black or bitwise or tolerance or impulses" 2>&1 | tail -2 ) | tee gpurun_out/r2l_pytest.log
