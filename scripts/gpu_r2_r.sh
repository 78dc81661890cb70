#!/bin/bash
# sanitizer runs of the sliding-window red-black kernels (the other kernels: gpu_r2_k.sh / profiles/r02_sanitizer_summary.txt)
OUT=gpurun_out
mkdir -p $OUT
run() {   # name tool timeout command...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to compute-sanitizer --tool $tool --print-limit 20 "$@" > $OUT/sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazards|passed|failed' $OUT/sanitizer_${name}_${tool}.log | tail -2 | tr '\n' ' ')"
}
for tool in memcheck initcheck synccheck racecheck; do
  EQ_RB_KERNEL=stream run rb_stream_256 $tool 600 python scripts/prof_frame.py tiny256 red_black 2
  EQ_RB_KERNEL=slide run rb_slide_256 $tool 600 python scripts/prof_frame.py tiny256 red_black 2
done 2>&1 | tee $OUT/sanitizer_summary_rbs.log
