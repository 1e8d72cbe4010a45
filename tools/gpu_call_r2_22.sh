#!/bin/bash
# GPU call 22: MFM backward fused into the Split firing, batched post-filter of the sub-tree maps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lightcnn.py tests/test_layerwise_subtree.py tests/test_inpaintgame.py tests/test_generic_sweeps.py -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2x_tests.log
timeout 300 python tools/generic_profile.py lightcnn > gpurun_out/r2x_profile_lightcnn.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2x_profile_weighted_subtree.log 2>&1
for w in weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2x_bench_$w.json 2> gpurun_out/r2x_bench_$w.err
done
tail -n 4 gpurun_out/r2x_tests.log | cut -c1-300
grep -A 14 "ms per call" gpurun_out/r2x_profile_lightcnn.log | cut -c1-170
grep -A 6 "ms per call" gpurun_out/r2x_profile_weighted_subtree.log | cut -c1-170
for w in weighted_subtree lightcnn; do cut -c1-200 gpurun_out/r2x_bench_$w.json; tail -n 2 gpurun_out/r2x_bench_$w.err | cut -c1-200; done
