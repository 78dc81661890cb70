#!/bin/bash
# Evidence for profiles/: ncu launch list of the bench command + one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench_c4.log 2>&1
tail -1 gpurun_out/ncu_bench_c4.log | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_tb -s 1 -c 1 \
   -o gpurun_out/prof_lsx_final_c4 -f python scripts/prof_linsolve.py 16384 20 2 1 > gpurun_out/ncu_lsx_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linsolve_tb -s 1 -c 1 \
   -o gpurun_out/prof_lsx_final_c3 -f python scripts/prof_linsolve.py 4096 40 2 1 > gpurun_out/ncu_lsx_c3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_rb_tiled -s 2 -c 1 \
   -o gpurun_out/prof_rb_final_c4 -f python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/ncu_rb_c4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_advect|k_divergence|k_gradient" -c 6 \
   -o gpurun_out/prof_stencils_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_st_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
