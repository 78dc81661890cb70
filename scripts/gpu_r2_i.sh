#!/bin/bash
# Round-2 call I (8 GPUs): multi-GPU parity, then bench.py on 2 / 8 GPUs with the default library and with the
# -DRBR_INLINE_BARRIER=1 build (k_rb_reg raises / waits for its neighbours' flags itself: no barrier kernel between launches)
N=${1:-8}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/r2i_pytest.log
for V in default ib; do
  if [ $V = ib ]; then export EQUILIBRIUM_CUDA_LIB=$PWD/variants/libeq_ib.so; else unset EQUILIBRIUM_CUDA_LIB; fi
  ( timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q -k "red_black" 2>&1 | tail -1 ) | sed "s/^/$V parity: /"
  for G in 2 $N; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$G \
        bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2i_${V}_n$G.json 2> gpurun_out/r2i_${V}_n$G.err
    python - "$V" "$G" <<'PY'
import json, sys
v, g = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/r2i_{v}_n{g}.json").read().strip().splitlines()[-1])
    rb, ex = d["red_black"], d["exact"]
    print(v, "N=", g, "exact ms", round(ex["ms_per_step"], 2), "rb ms", round(rb["ms_per_step"], 3), "rb phases",
          {k: round(x, 3) for k, x in rb["roofline"]["phases_ms_per_step"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
except Exception as e:
    print(v, g, "failed", e)
PY
  done
done 2>&1 | tee gpurun_out/r2i.log
