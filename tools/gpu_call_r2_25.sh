#!/bin/bash
# GPU call 25: programmatic dependent launch at batch 256 (A = XFRB_PDL=1, B = default), twice each
mkdir -p gpurun_out
for v in A B A B; do
  if [ $v = A ]; then export XFRB_PDL=1; else unset XFRB_PDL; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2ab_bench_$v.json 2> gpurun_out/r2ab_bench_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r2ab_bench_$v.json'))
print('variant $v', round(d['value']), 'e2e', round(d['e2e']['value']), 'bwd', round(d['roofline']['bwd_ms_per_step'], 2), d['clocks']['sm_mhz'])
PY
done
