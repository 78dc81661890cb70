#!/bin/bash
# Round-2 call B (one GPU): stage isolation of k_linsolve_tb (variants built locally with -DEQ_DEBUG_KNOBS; WRONG results on
# purpose, dependency waits off) -- which warp role sets the per-job rate?
mkdir -p gpurun_out
{
python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/default coupled: /"
for V in dbg dbg_NOLOAD dbg_NOSTORE dbg_NOCOMPUTE dbg_NOLS dbg_NOCS; do
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_$V.so python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/$V nodeps: /"
done
for C in 1 2 3 4; do
  EQ_LSX_CTAS_PER_SM=$C EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg.so python scripts/prof_linsolve.py 16384 20 2 2 | tail -1 | sed "s/^/dbg nodeps ctas_per_sm=$C: /"
done
} 2>&1 | tee gpurun_out/r2b.log
