#!/bin/bash
# session 3, run B: float4 stencils, single-pass advect, register-tiled red-black -- parity, timings, ncu
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/s3b.log
for O in 0 1 2; do
  timeout 300 python scripts/prof_linsolve.py 16384 20 $O 3 red_black 2>&1 | tail -1 | sed "s/^/reg /"
done | tee -a gpurun_out/s3b.log
EQ_RB_KERNEL=tiled timeout 300 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/tiled /" | tee -a gpurun_out/s3b.log
timeout 300 python scripts/prof_linsolve.py 4096 40 2 3 red_black 2>&1 | tail -1 | sed "s/^/reg /" | tee -a gpurun_out/s3b.log
timeout 300 python scripts/prof_linsolve.py 4096 40 0 3 red_black 2>&1 | tail -1 | sed "s/^/reg /" | tee -a gpurun_out/s3b.log
timeout 900 python bench.py --workload c4 --steps 3 > gpurun_out/s3b_bench_c4.json 2> gpurun_out/s3b_bench_c4.err
python - <<'PY' | tee -a gpurun_out/s3b.log
import json
for w in ["c4"]:
    try:
        d = json.loads(open(f"gpurun_out/s3b_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"))
    except Exception as e:
        print(w, "failed", e)
PY
timeout 600 python bench.py --workload c3 --steps 5 > gpurun_out/s3b_bench_c3.json 2> gpurun_out/s3b_bench_c3.err
python - <<'PY' | tee -a gpurun_out/s3b.log
import json
for w in ["c3"]:
    try:
        d = json.loads(open(f"gpurun_out/s3b_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "ms/step", d["ms_per_step"], "phases", d["roofline"]["phases_ms_per_step"], "rb", d.get("red_black"), "e2e", d.get("e2e"))
    except Exception as e:
        print(w, "failed", e)
PY
# launch list (per-kernel durations) of the c3 bench command, and full captures of the stencil / red-black kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3b_launches_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-extras > gpurun_out/s3b_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_advect|k_divergence|k_gradient" -s 8 -c 4 \
   -o gpurun_out/s3b_stencils_c3 -f python bench.py --workload c3 --steps 1 --warmup 3 --no-extras > gpurun_out/s3b_ncu_st_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_reg -s 3 -c 1 \
   -o gpurun_out/s3b_rbreg_c3 -f python scripts/prof_linsolve.py 4096 40 2 1 red_black > gpurun_out/s3b_ncu_rb_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb_reg -s 3 -c 1 \
   -o gpurun_out/s3b_rbreg_c4 -f python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/s3b_ncu_rb_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tee -a gpurun_out/s3b.log
