#!/bin/bash
# Round-2 call A (one GPU): the driver's exact test command first, with the full log kept; then baselines.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_smi.log 2>&1
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
python - <<'PY' | tee -a gpurun_out/r2a.log
import json
for name in ("c4", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/r2a_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "exact ms", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "rb ms",
              round(d["red_black"]["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "clocks", d["clocks"])
    except Exception as e:
        print(name, "failed", e)
PY
