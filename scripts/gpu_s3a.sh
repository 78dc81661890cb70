#!/bin/bash
# session 3, run A: phase-structured fast loop (interleaved sub-step chains) -- parity + timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/s3a.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee -a gpurun_out/s3a.log
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=2 /"
done | tee -a gpurun_out/s3a.log
EQ_LSX_NODEPS=1 timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=2 nodeps /" | tee -a gpurun_out/s3a.log
for PB in 1 2 8; do
EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=2 pub=$PB /" | tee -a gpurun_out/s3a.log
done
timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 2>&1 | tail -1 | sed "s/^/T=2 /" | tee -a gpurun_out/s3a.log
timeout 300 python scripts/prof_linsolve.py 4096 40 0 3 2>&1 | tail -1 | sed "s/^/T=2 /" | tee -a gpurun_out/s3a.log
for T in 3 4; do
  for O in 0 2; do
    EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 2>&1 | tail -1 | sed "s/^/T=$T /"
  done
  EQ_LSX_NODEPS=1 EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/T=$T nodeps /"
  EQUILIBRIUM_CUDA_LIB=variants/libeq_T$T.so timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 2>&1 | tail -1 | sed "s/^/T=$T /"
done | tee -a gpurun_out/s3a.log
timeout 600 python bench.py --workload c4 --no-extras --steps 3 > gpurun_out/s3a_bench_c4.json 2> gpurun_out/s3a_bench_c4.err
cut -c1-300 gpurun_out/s3a_bench_c4.json | tee -a gpurun_out/s3a.log
timeout 600 python bench.py --workload c3 --no-extras --steps 5 > gpurun_out/s3a_bench_c3.json 2> gpurun_out/s3a_bench_c3.err
cut -c1-300 gpurun_out/s3a_bench_c3.json | tee -a gpurun_out/s3a.log
