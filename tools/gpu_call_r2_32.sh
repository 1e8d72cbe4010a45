#!/bin/bash
# GPU call 32: final validation - full suite, smoke, the default bench line in its driver form
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2ai_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2ai_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ai_smoke.log 2>&1
XFRB_BENCH_LAUNCHES=gpurun_out/r2ai_launches.jsonl timeout 600 python bench.py > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r2ai_bench_reference.json 2> gpurun_out/r2ai_bench_reference.err
grep -v "^$" gpurun_out/r2ai_pytest.log | tail -n 6 | cut -c1-300
tail -n 2 gpurun_out/r2ai_smoke.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ai_bench.json'))
print(round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'bwd', round(d['roofline']['bwd_ms_per_step'], 2), d['clocks'], d.get('latency_ms_batch1'), d.get('cpu_baseline', {}).get('value'), d['gpu_launches'])
r = json.load(open('gpurun_out/r2ai_bench_reference.json'))
print('reference arm', r['value'], r['cpu_baseline']['cores'])
PY
python tools/launch_roofline.py gpurun_out/r2ai_launches.jsonl 2232 2 | head -12
