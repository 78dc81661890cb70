#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -c 3000 gpurun_out/bench_c4.json
timeout 600 python bench.py --workload c3 --no-extras > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cut -c1-400 gpurun_out/bench_c3.json
scripts/gpu_profiles.sh
