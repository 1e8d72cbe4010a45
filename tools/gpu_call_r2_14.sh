#!/bin/bash
# GPU call 14: vectorised xfrb_hook + device prior tables + graph-replayed generic sweeps + lc_conv1 rewrite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2o_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2o_pytest.log
for w in layer_sweep weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2o_bench_$w.json 2> gpurun_out/r2o_bench_$w.err
done
timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
grep -v "^$" gpurun_out/r2o_pytest.log | tail -8 | cut -c1-300
for w in layer_sweep weighted_subtree lightcnn; do cut -c1-260 gpurun_out/r2o_bench_$w.json; tail -n 3 gpurun_out/r2o_bench_$w.err | cut -c1-300; done
cut -c1-200 gpurun_out/r2o_bench.json; tail -n 2 gpurun_out/r2o_bench.err
