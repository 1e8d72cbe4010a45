#!/bin/bash
# GPU call 9: restored epilogue layout; whole suite; default bench with latency (CUDA graphs) + comparator + cpu baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2i_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2i_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
grep -v "^$" gpurun_out/r2i_pytest.log | tail -n 25 | cut -c1-250
tail -n 2 gpurun_out/r2i_smoke.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('latency_ms_batch1'), d.get('gpu_library_baseline'), d.get('cpu_baseline'))
PY
tail -n 3 gpurun_out/r2i_bench.err
