#!/bin/bash
# Round-2 call J (one GPU): the driver's commands on the final tree -- tests, smoke, default bench, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2j_pytest.log 2>&1
tail -4 gpurun_out/r2j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py ) > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -3 gpurun_out/r2j_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; tail -3 gpurun_out/r2j_ref.err
