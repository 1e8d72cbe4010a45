#!/bin/bash
# GPU call 30: steady-state records of the generic workloads with the row-walk hook kernel
mkdir -p gpurun_out
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2ag_profile_layer_sweep.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2ag_profile_weighted_subtree.log 2>&1
for w in layer_sweep weighted_subtree; do
  timeout 400 python bench.py --workload $w --warmup 6 --no-cpu-baseline > gpurun_out/r2ag_bench_$w.json 2> gpurun_out/r2ag_bench_$w.err
done
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2ag_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2ag_pytest.log
for f in layer_sweep weighted_subtree; do grep -v Warn gpurun_out/r2ag_profile_$f.log | head -9 | cut -c1-170; done
for w in layer_sweep weighted_subtree; do cut -c1-200 gpurun_out/r2ag_bench_$w.json; tail -n 2 gpurun_out/r2ag_bench_$w.err | cut -c1-200; done
grep -v "^$" gpurun_out/r2ag_pytest.log | tail -n 6 | cut -c1-300
