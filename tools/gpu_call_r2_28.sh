#!/bin/bash
# GPU call 28: row-walk hook kernel (A = default on, B = XFRB_HOOK_ROWS=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generic_sweeps.py tests/test_layerwise_subtree.py tests/test_inpaintgame.py tests/test_resnet50_128.py -m gpu -q -x > gpurun_out/r2ae_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2ae_tests.log
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2ae_profile_weighted_subtree_A.log 2>&1
XFRB_HOOK_ROWS=0 timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2ae_profile_weighted_subtree_B.log 2>&1
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2ae_profile_layer_sweep_A.log 2>&1
XFRB_HOOK_ROWS=0 timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2ae_profile_layer_sweep_B.log 2>&1
tail -n 4 gpurun_out/r2ae_tests.log | cut -c1-300
for f in weighted_subtree_A weighted_subtree_B layer_sweep_A layer_sweep_B; do echo $f; grep -A 9 "ms per call" gpurun_out/r2ae_profile_$f.log | grep -v Warn | cut -c1-150; done
