#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) | tee gpurun_out/r2s_pytest.log
for w in c1 c2; do
  timeout 600 python bench.py --workload $w --no-configs > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err
  python - $w <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2s_bench_{sys.argv[1]}.json').read().strip().split('\n')[-1])
print(sys.argv[1], 'rb', d['value'], d['ms_per_step'], 'exact', d['exact']['value'], d['exact']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['roofline']['phases_ms_per_step'])
PY
done
