#!/bin/bash
# GPU call 35: final build - full suite, smoke, default bench line in its driver form
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2an_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2an_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2an_smoke.log 2>&1
XFRB_BENCH_LAUNCHES=gpurun_out/r2an_launches.jsonl timeout 600 python bench.py > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err
grep -v "^$" gpurun_out/r2an_pytest.log | tail -n 5 | cut -c1-300
tail -n 2 gpurun_out/r2an_smoke.log
wc -l gpurun_out/r2an_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2an_bench.json'))
print(round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'bwd', round(d['roofline']['bwd_ms_per_step'], 2), d['clocks'], d.get('latency_ms_batch1'), d.get('cpu_baseline', {}).get('value'), d['gpu_launches'])
PY
