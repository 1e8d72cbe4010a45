#!/bin/bash
# GPU call 16: hook chains + row skipping, multi-block score / threshold kernels, arg-byte maxpool backward, join / stem_bwd / contrast tuning
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2q_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2q_pytest.log
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2q_profile_layer_sweep.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2q_profile_weighted_subtree.log 2>&1
XFRB_BENCH_LAUNCHES=gpurun_out/r2q_launches.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
timeout 300 python bench.py --workload lightcnn --no-cpu-baseline > gpurun_out/r2q_bench_lightcnn.json 2> gpurun_out/r2q_bench_lightcnn.err
grep -v "^$" gpurun_out/r2q_pytest.log | tail -n 12 | cut -c1-300
tail -n 22 gpurun_out/r2q_profile_layer_sweep.log | cut -c1-200
tail -n 22 gpurun_out/r2q_profile_weighted_subtree.log | cut -c1-200
python - <<'PY'
import json
for f in ('r2q_bench', 'r2q_bench_lightcnn'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, d['value'], d['e2e'], d['roofline']['frac'], d['roofline'].get('bwd_ms_per_step'), d['clocks'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 gpurun_out/r2q_bench.err gpurun_out/r2q_bench_lightcnn.err
