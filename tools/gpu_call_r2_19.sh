#!/bin/bash
# GPU call 19: records with the final defaults (pairs on small problems, narrow dgrad tiles, pinned ebp_batch results)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2t_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2t_pytest.log
XFRB_BENCH_LAUNCHES=gpurun_out/r2t_launches.jsonl timeout 600 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
for w in layer_sweep weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w > gpurun_out/r2t_bench_$w.json 2> gpurun_out/r2t_bench_$w.err
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1
grep -v "^$" gpurun_out/r2t_pytest.log | tail -n 6 | cut -c1-300
python - <<'PY'
import json
for f in ('r2t_bench', 'r2t_bench_layer_sweep', 'r2t_bench_weighted_subtree', 'r2t_bench_lightcnn'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'frac', round(d['roofline']['frac'], 3), d['roofline'].get('bwd_ms_per_step'), d['clocks'], d.get('latency_ms_batch1'), d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 2 gpurun_out/r2t_bench.err gpurun_out/r2t_smoke.log
