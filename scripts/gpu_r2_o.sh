#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2o_pytest.log
timeout 600 python scripts/measure_traffic.py > gpurun_out/r2o_traffic.log 2>&1; tail -3 gpurun_out/r2o_traffic.log
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 400 gpurun_out/r2o_bench.json
