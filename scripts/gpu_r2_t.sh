#!/bin/bash
# small-slab study on one GPU (timing only: EQ_RQ_DEBUG_ROWS makes k_rb_stream work on the first R rows of the grid, as a
# rank of an 8- or 4-GPU run would): which segment count suits a slab of 2048 / 4096 rows x 16384 columns?
mkdir -p gpurun_out
export EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg.so
{
for rows in 2048 4096; do
  for segs in 7 11 14 16 21 28 32 43 64; do
    EQ_RQ_DEBUG_ROWS=$rows EQ_RQ_SEGS=$segs timeout 120 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/rows=$rows segs=$segs: /"
  done
done
} 2>&1 | tee gpurun_out/r2t.log
