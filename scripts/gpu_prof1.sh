#!/bin/bash
mkdir -p gpurun_out
for args in "4096 1 2" "4096 4 2" "4096 40 2" "4096 40 0" "4096 40 1" "16384 1 2" "16384 20 2" "1024 20 2"; do
  timeout 120 python scripts/prof_linsolve.py $args 3
done 2>&1 | tee gpurun_out/linsolve_times.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_exact -s 1 -c 1 \
   -o gpurun_out/prof_lsx_r1 -f python scripts/prof_linsolve.py 4096 40 0 1 > gpurun_out/ncu_lsx.log 2>&1
tail -3 gpurun_out/ncu_lsx.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_advect_nan_and_huge_velocities 2>&1 | tail -15
