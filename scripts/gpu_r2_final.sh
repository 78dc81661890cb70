#!/bin/bash
# the driver's commands, in its order, plus the traffic record for the final sources
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2final_pytest.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/r2final_smoke.log
timeout 600 python scripts/measure_traffic.py > gpurun_out/r2final_traffic.log 2>&1; tail -2 gpurun_out/r2final_traffic.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; tail -c 300 gpurun_out/r2final_bench.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_ref.json 2> gpurun_out/r2final_ref.err; tail -c 600 gpurun_out/r2final_ref.json
