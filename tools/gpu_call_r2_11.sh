#!/bin/bash
# GPU call 11: 32-bit index arithmetic in the elementwise kernels; suite + all four workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2k_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2k_pytest.log
XFRB_BENCH_LAUNCHES=gpurun_out/r2k_launches.jsonl timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
for w in layer_sweep weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_bench_$w.json 2> gpurun_out/r2k_bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_ncu_launches.csv python bench.py --no-cpu-baseline --no-extras --batch 128 --chunk 128 --steps 1 --warmup 3 > gpurun_out/r2k_ncu_b.log 2>&1
grep -v "^$" gpurun_out/r2k_pytest.log | tail -n 8 | cut -c1-250
for f in r2k_bench r2k_bench_layer_sweep r2k_bench_weighted_subtree r2k_bench_lightcnn; do cut -c1-150 gpurun_out/$f.json; tail -n 2 gpurun_out/$f.err | cut -c1-200; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['bwd_ms_per_step'], d.get('latency_ms_batch1'))
PY
