#!/bin/bash
# Round-2 starter: which stage of the k_linsolve_tb job pipeline sets the per-job rate?  Builds stage-isolation variants
# (results are WRONG on purpose, dependency waits off) and times a 16384^2, K=20 Passive solve with each.
#   base nodeps            : full pipeline, no dependency waits            (12.1 ms in round 1)
#   NOLOAD / NOSTORE / NOCOMPUTE : the loader / storer / compute warp only signal
#   SPLIT + the same       : one compute warp per sub-step
# Run under gpurun on one GPU:  scripts/gpu_r2_limits.sh
set -e
mkdir -p gpurun_out variants
FLAGS="-DEQ_DEBUG_KNOBS -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
for D in NOLOAD NOSTORE NOCOMPUTE; do
  [ -f variants/libeq_dbg_$D.so ] || nvcc $FLAGS -DTBX_DBG_$D -o variants/libeq_dbg_$D.so equilibrium_b200/csrc/eq_api.cu
  [ -f variants/libeq_dbg_SPLIT_$D.so ] || nvcc $FLAGS -DTBX_SPLIT=1 -DTBX_DBG_$D -o variants/libeq_dbg_SPLIT_$D.so equilibrium_b200/csrc/eq_api.cu
done
[ -f variants/libeq_SPLIT.so ] || nvcc $FLAGS -DTBX_SPLIT=1 -o variants/libeq_SPLIT.so equilibrium_b200/csrc/eq_api.cu
{
EQ_LSX_NODEPS=1 python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/base nodeps /"
EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/SPLIT nodeps /"
for D in NOLOAD NOSTORE NOCOMPUTE; do
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg_$D.so python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/$D nodeps /"
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_dbg_SPLIT_$D.so python scripts/prof_linsolve.py 16384 20 2 3 | tail -1 | sed "s/^/SPLIT $D nodeps /"
done
} | tee gpurun_out/r2_limits.log
