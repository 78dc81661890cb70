#!/bin/bash
# Round-2 first call (one GPU): everything that round 1 left unverified on hardware, then fresh baselines.
#   1. the whole -m gpu suite (the last three tests -- golden device noise, the C++ host mirror, the Python
#      CurrentSimulation loop -- have only run on the emulated build so far)
#   2. bench.py on config 4 and config 3 (baselines for the round) and the reference arm
#   3. launch list of the bench command for profiles/
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/r2_first.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu_bench_c4.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_first.log
import json
for name in ("c4", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "exact ms", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "rb ms",
              round(d["red_black"]["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "clocks", d["clocks"])
    except Exception as e:
        print(name, "failed", e)
PY
