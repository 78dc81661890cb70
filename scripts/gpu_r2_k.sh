#!/bin/bash
# Round-2 call K (8 GPUs): multi-GPU parity on the final tree, then the driver's scaling series (bench.py on 2 / 4 / 8 GPUs)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/r2k_pytest.log
for G in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$G \
      bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2k_n$G.json 2> gpurun_out/r2k_n$G.err
  python - "$G" <<'PY'
import json, sys
g = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2k_n{g}.json").read().strip().splitlines()[-1])
    rb, ex = d["red_black"], d["exact"]
    print("N=", g, "value", round(d["value"] / 1e9, 2), "G/s", d["headline_mode"], "ms", round(d["ms_per_step"], 3), "| exact ms", round(ex["ms_per_step"], 2),
          "| rb phases", {k: round(x, 3) for k, x in rb["roofline"]["phases_ms_per_step"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
except Exception as e:
    print(g, "failed", e)
PY
done 2>&1 | tee gpurun_out/r2k.log
