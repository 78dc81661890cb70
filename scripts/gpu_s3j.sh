#!/bin/bash
# session 3, run J: one compute warp per sub-step (TBX_SPLIT) against the fused single compute warp; 8-warp red-black CTAs
mkdir -p gpurun_out
( EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 ) | sed "s/^/SPLIT /" | tee gpurun_out/s3j.log
for O in 0 1 2; do
  EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/SPLIT /"
done | tee -a gpurun_out/s3j.log
EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/SPLIT nodeps /" | tee -a gpurun_out/s3j.log
for PB in 2 8; do
EQ_LSX_PUBBATCH=$PB EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/SPLIT pub=$PB /" | tee -a gpurun_out/s3j.log
done
for O in 0 2; do
EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 300 python scripts/prof_linsolve.py 4096 40 $O 3 2>&1 | tail -1 | sed "s/^/SPLIT /" | tee -a gpurun_out/s3j.log
done
for PB in 1 4; do
EQ_LSX_PUBBATCH=$PB EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 2>&1 | tail -1 | sed "s/^/SPLIT pub=$PB /" | tee -a gpurun_out/s3j.log
done
timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/base /" | tee -a gpurun_out/s3j.log
for O in 0 2; do
EQUILIBRIUM_CUDA_LIB=variants/libeq_W8.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 red_black 2>&1 | tail -1 | sed "s/^/W8 /" | tee -a gpurun_out/s3j.log
done
EQUILIBRIUM_CUDA_LIB=variants/libeq_W8.so timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 red_black 2>&1 | tail -1 | sed "s/^/W8 /" | tee -a gpurun_out/s3j.log
EQUILIBRIUM_CUDA_LIB=variants/libeq_SPLIT.so timeout 600 python bench.py --workload c4 --steps 3 --no-extras > gpurun_out/s3j_bench_c4_split.json 2> gpurun_out/s3j_bench_c4_split.err
cut -c1-260 gpurun_out/s3j_bench_c4_split.json | tee -a gpurun_out/s3j.log
