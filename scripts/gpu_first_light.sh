#!/bin/bash
# First GPU pass: smoke, parity tests, bench lines, ncu launch list.  Run under gpurun.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 -x 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench c3" ; timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c3.json
echo "== bench c4" ; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c4.json
echo "== ncu launch list (c3, exact)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
