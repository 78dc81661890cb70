#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_red_black.py -m gpu -q -x 2>&1 | tail -3
for args in "16384 20 2 3 red_black" "16384 20 1 3 red_black" "4096 40 2 3 red_black" "1024 20 2 3 red_black"; do timeout 120 python scripts/prof_linsolve.py $args; done 2>&1 | tee gpurun_out/rb_times.txt
