#!/bin/bash
# compute-sanitizer over small runs of every wavefront / tile kernel: memcheck, initcheck, racecheck, synccheck.
# usage (one GPU): scripts/sanitize.sh [outdir]
OUT=${1:-gpurun_out}
mkdir -p $OUT
run() {   # name tool timeout command...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to compute-sanitizer --tool $tool --print-limit 20 "$@" > $OUT/sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazards|passed|failed' $OUT/sanitizer_${name}_${tool}.log | tail -2 | tr '\n' ' ')"
}

for tool in memcheck initcheck synccheck racecheck; do
  run smoke_exact $tool 600 python -c "import __graft_entry__ as g; g.smoke()"
  run rb_256 $tool 600 python scripts/prof_frame.py tiny256 red_black 2
  EQ_EXACT_KERNEL=wf run wf_256 $tool 600 python scripts/prof_frame.py tiny256 exact 2
done 2>&1 | tee $OUT/sanitizer_summary.log
