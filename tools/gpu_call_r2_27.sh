#!/bin/bash
# GPU call 27: JOIN on 128-wide tiles (A) vs default (B); graph-replayed layer priors; tests
mkdir -p gpurun_out
for v in A B; do
  unset XFRB_JOIN_BN
  if [ $v = A ]; then export XFRB_JOIN_BN=128; fi
  XFRB_BENCH_LAUNCHES=gpurun_out/r2ad_launches_$v.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2ad_bench_$v.json 2> gpurun_out/r2ad_bench_$v.err
  echo "VARIANT $v"; python tools/launch_roofline.py gpurun_out/r2ad_launches_$v.jsonl 2232 2 2>/dev/null | grep "dgrad_join" | head -3; cut -c1-120 gpurun_out/r2ad_bench_$v.json
done
unset XFRB_JOIN_BN
timeout 900 python -m pytest tests/test_layerwise_subtree.py tests/test_generic_sweeps.py tests/test_bf16x2.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2ad_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2ad_tests.log
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2ad_profile_layer_sweep.log 2>&1
timeout 400 python bench.py --workload layer_sweep --no-cpu-baseline > gpurun_out/r2ad_bench_layer_sweep.json 2> gpurun_out/r2ad_bench_layer_sweep.err
tail -n 3 gpurun_out/r2ad_tests.log | cut -c1-300
grep -A 9 "ms per call" gpurun_out/r2ad_profile_layer_sweep.log | cut -c1-170
cut -c1-200 gpurun_out/r2ad_bench_layer_sweep.json; tail -n 2 gpurun_out/r2ad_bench_layer_sweep.err | cut -c1-200
