#!/bin/bash
# session 3, run L: final state of the round -- GPU tests, smoke, default bench lines (c4, c3), reference arm
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/s3l.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee -a gpurun_out/s3l.log
timeout 900 python bench.py > gpurun_out/s3l_bench_c4.json 2> gpurun_out/s3l_bench_c4.err
tail -2 gpurun_out/s3l_bench_c4.err
timeout 600 python bench.py --workload c3 > gpurun_out/s3l_bench_c3.json 2> gpurun_out/s3l_bench_c3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3l_bench_ref.json 2> gpurun_out/s3l_bench_ref.err
python - <<'PY' | tee -a gpurun_out/s3l.log
import json
for w in ["c4", "c3"]:
    try:
        d = json.loads(open(f"gpurun_out/s3l_bench_{w}.json").read().strip().splitlines()[-1])
        rb = d.get("red_black") or {}
        print(w, "exact ms", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "roofline frac %.3f" % d["roofline"]["frac"], "phases", {k: round(v, 3) for k, v in d["roofline"]["phases_ms_per_step"].items()},
              "| rb ms", round(rb.get("ms_per_step", 0), 3), "rb frac %.3f" % rb["roofline"]["frac"], "phys", rb["roofline"]["physical_frac"],
              "| e2e ms", round(d["e2e"]["ms_per_step"], 3), "mirror ms", round(d["e2e_full_mirror"]["ms_per_step"], 3), "| cpu %.4g" % d["cpu_baseline"]["value"], "clocks", d["clocks"], "launches", d["gpu_launches"])
    except Exception as e:
        print(w, "failed", e)
print(open("gpurun_out/s3l_bench_ref.json").read()[:300])
PY
