#!/bin/bash
# GPU call 23: pair-tensor dgrads in the firing-by-firing sweeps (bf16x2 plan); full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2z_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2z_pytest.log
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2z_profile_layer_sweep.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2z_profile_weighted_subtree.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree tf32x3 > gpurun_out/r2z_profile_weighted_subtree_tf32.log 2>&1
for w in layer_sweep weighted_subtree; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2z_bench_$w.json 2> gpurun_out/r2z_bench_$w.err
done
grep -v "^$" gpurun_out/r2z_pytest.log | tail -n 12 | cut -c1-300
grep -A 9 "ms per call" gpurun_out/r2z_profile_layer_sweep.log | cut -c1-170
grep -A 9 "ms per call" gpurun_out/r2z_profile_weighted_subtree.log | cut -c1-170
grep -A 2 "ms per call" gpurun_out/r2z_profile_weighted_subtree_tf32.log | cut -c1-170
for w in layer_sweep weighted_subtree; do cut -c1-200 gpurun_out/r2z_bench_$w.json; tail -n 2 gpurun_out/r2z_bench_$w.err | cut -c1-200; done
