#!/bin/bash
# Round-2 call H (one GPU): driver's test command, traffic record, the default bench line, sanitizer summaries
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
timeout 600 python scripts/measure_traffic.py > gpurun_out/r2h_traffic.log 2>&1; tail -2 gpurun_out/r2h_traffic.log
( time timeout 900 python bench.py ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err
timeout 900 scripts/sanitize.sh gpurun_out
