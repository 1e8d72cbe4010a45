#!/bin/bash
# GPU call 24: sweep size (probes per engine sweep) against L2 residency of the inter-kernel tensors
mkdir -p gpurun_out
for c in 256 128 64 32; do
  timeout 300 python bench.py --no-cpu-baseline --no-extras --chunk $c > gpurun_out/r2aa_bench_chunk$c.json 2> gpurun_out/r2aa_bench_chunk$c.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r2aa_bench_chunk$c.json'))
print('chunk $c', round(d['value']), 'e2e', round(d['e2e']['value']), 'bwd', round(d['roofline']['bwd_ms_per_step'], 2), d['clocks']['sm_mhz'], [(k['launches'], round(k['avg_us'], 1)) for k in d['roofline']['kernels']])
PY
done
