#!/bin/bash
# session 3, run E: branch-free code staging in k_rb_reg, 32-bit advect offsets, final-ish bench + profiles
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/s3e.log
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 red_black 2>&1 | tail -1 | sed "s/^/reg /"
done | tee -a gpurun_out/s3e.log
for O in 0 1 2; do
timeout 300 python scripts/prof_linsolve.py 4096 40 $O 3 red_black 2>&1 | tail -1 | sed "s/^/reg /" | tee -a gpurun_out/s3e.log
done
for W in c4 c3; do
timeout 900 python bench.py --workload $W > gpurun_out/s3e_bench_$W.json 2> gpurun_out/s3e_bench_$W.err
tail -3 gpurun_out/s3e_bench_$W.err
python - $W <<'PY' | tee -a gpurun_out/s3e.log
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/s3e_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, "ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"), "mirror", d.get("e2e_full_mirror"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print(w, "failed", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_advect|k_divergence|k_gradient" -s 8 -c 4 \
   -o gpurun_out/s3e_stencils_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/s3e_ncu_st_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_advect|k_divergence|k_gradient" -s 8 -c 4 \
   -o gpurun_out/s3e_stencils_c3 -f python bench.py --workload c3 --steps 1 --warmup 3 --no-extras > gpurun_out/s3e_ncu_st_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_reg -s 3 -c 1 \
   -o gpurun_out/s3e_rbreg_c4 -f python scripts/prof_linsolve.py 16384 20 0 1 red_black > gpurun_out/s3e_ncu_rb_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_reg -s 3 -c 1 \
   -o gpurun_out/s3e_rbreg_c3 -f python scripts/prof_linsolve.py 4096 40 0 1 red_black > gpurun_out/s3e_ncu_rb_c3.log 2>&1
ls -la gpurun_out/s3e*.ncu-rep | tee -a gpurun_out/s3e.log
