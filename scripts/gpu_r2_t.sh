#!/bin/bash
# small-slab / mid-size-grid study on one GPU (timing only: EQ_RQ_DEBUG_ROWS makes k_rb_stream work on the first R rows of
# the grid, as a rank of an 8- or 4-GPU run would): interior segment count and rows per wall-strip task
mkdir -p gpurun_out
export EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg.so
{
for rows in 2048 4096 8192; do
  EQ_RQ_DEBUG_ROWS=$rows timeout 120 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/rows=$rows default: /"
done
for segs in 0 21 28 42; do
  for o in 2 0; do
    if [ $segs = 0 ]; then unset EQ_RQ_SEGS; else export EQ_RQ_SEGS=$segs; fi
    timeout 120 python scripts/prof_linsolve.py 4096 40 $o 3 red_black 2>&1 | tail -1 | sed "s/^/c3 segs=$segs: /"
  done
done
unset EQ_RQ_SEGS
for n in 2048 3072 8192; do timeout 120 python scripts/prof_linsolve.py $n 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/default: /"; done
for n in 2048 3072; do EQ_RB_KERNEL=reg timeout 120 python scripts/prof_linsolve.py $n 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/k_rb_reg: /"; done
for n in 1024 1536; do EQ_RB_KERNEL=stream timeout 120 python scripts/prof_linsolve.py $n 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/stream forced: /"; done
for n in 1024 1536; do timeout 120 python scripts/prof_linsolve.py $n 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/default (reg): /"; done
for o in 0 2; do timeout 120 python scripts/prof_linsolve.py 16384 20 $o 3 red_black 2>&1 | tail -1 | sed "s/^/full grid default: /"; done
} 2>&1 | tee gpurun_out/r2t3.log
