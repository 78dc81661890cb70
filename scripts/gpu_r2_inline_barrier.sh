#!/bin/bash
# Round-2 starter (N GPUs, default 8: gpurun --gpus 8 -- scripts/gpu_r2_inline_barrier.sh 8): A/B of -DRBR_INLINE_BARRIER=1
# (k_rb_reg waits for / raises its neighbours' flags itself; no barrier kernel between the launches of a row-slab solve)
# against the default build.  Parity first (multi-GPU red-black test through each library), then bench.py's red_black
# object at N GPUs.  Round 1: default 10.3 ms/frame at N=8 (5.96x of one GPU); the variant is verified on the emulator only.
N=${1:-8}
set -e
mkdir -p gpurun_out variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
[ -f variants/libeq_ib.so ] || nvcc $FLAGS -DRBR_INLINE_BARRIER=1 -o variants/libeq_ib.so equilibrium_b200/csrc/eq_api.cu
set +e
for V in default ib default ib; do
  if [ $V = ib ]; then export EQUILIBRIUM_CUDA_LIB=$PWD/variants/libeq_ib.so; else unset EQUILIBRIUM_CUDA_LIB; fi
  ( timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q -k "red_black" 2>&1 | tail -2 ) | sed "s/^/$V parity: /"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_ib_$V.json 2> gpurun_out/r2_ib_$V.err
  python - "$V" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_ib_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    rb = d["red_black"]
    print(sys.argv[1], "rb ms/frame", round(rb["ms_per_step"], 3), "phases", {k: round(v, 3) for k, v in rb["phases_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done 2>&1 | tee gpurun_out/r2_inline_barrier.log
