#!/bin/bash
# GPU call 33: the 256-probe sweep of the e2e leg as a graph replay (A: XFRB_GRAPH_MAX_N=256) against eager launches (B)
mkdir -p gpurun_out
for v in A B A B; do
  if [ $v = A ]; then export XFRB_GRAPH_MAX_N=256; else unset XFRB_GRAPH_MAX_N; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 8 > gpurun_out/r2al_bench_$v.json 2> gpurun_out/r2al_bench_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r2al_bench_$v.json'))
print('variant $v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['clocks']['sm_mhz'])
PY
  tail -n 1 gpurun_out/r2al_bench_$v.err | cut -c1-200
done
