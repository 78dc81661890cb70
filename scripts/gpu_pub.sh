#!/bin/bash
mkdir -p gpurun_out
for PB in 1 2 3 4 6 8; do
  EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/tb pub=$PB /"
done | tee gpurun_out/pub_times.log
for PB in 1 2 4; do
  EQ_LSX_TB=0 EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 2>&1 | tail -1 | sed "s/^/v7 pub=$PB /"
  EQ_LSX_TB=0 EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | sed "s/^/v7 pub=$PB /"
  EQ_LSX_PUBBATCH=$PB timeout 300 python scripts/prof_linsolve.py 4096 20 2 3 2>&1 | tail -1 | sed "s/^/tb pub=$PB /"
done | tee -a gpurun_out/pub_times.log
