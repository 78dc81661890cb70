#!/bin/bash
mkdir -p gpurun_out
{
( EQ_RB_KERNEL=stream timeout 600 python -m pytest tests/test_red_black.py tests/test_gpu_parity.py -x -q -m gpu -k "red_black or bitwise or impulses" 2>&1 | tail -3 )
for cfg in "16384 20 2" "16384 20 0" "16384 20 1" "4096 40 2" "4096 40 0"; do
  set -- $cfg
  EQ_RB_KERNEL=stream timeout 120 python scripts/prof_linsolve.py $1 $2 $3 3 red_black 2>&1 | tail -1 | sed "s/^/stream $cfg: /"
done
for segs in 7 14 28 56 112; do
  EQ_RQ_SEGS=$segs EQ_RB_KERNEL=stream timeout 120 python scripts/prof_linsolve.py 16384 20 2 3 red_black 2>&1 | tail -1 | sed "s/^/stream segs=$segs: /"
done
for segs in 3 7 14 28; do
  EQ_RQ_SEGS=$segs EQ_RB_KERNEL=stream timeout 120 python scripts/prof_linsolve.py 4096 40 2 3 red_black 2>&1 | tail -1 | sed "s/^/stream c3 segs=$segs: /"
done
} 2>&1 | tee gpurun_out/r2m.log
EQ_RB_KERNEL=stream timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rb_stream -s 3 -c 1 -o gpurun_out/rq_stream -f python scripts/prof_linsolve.py 16384 20 2 1 red_black > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
