#!/bin/bash
# GPU call 18: optimised hook kernel (2-D grid, mode template), narrow tiles for small dgrads, CTA pairs on small problems
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2s_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2s_pytest.log
timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2s_batch1.log 2>&1
XFRB_SMALL_BN=0 timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2s_batch1_bn256.log 2>&1
XFRB_PAIR_MIN=1 timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2s_batch1_pairs.log 2>&1
XFRB_PAIR_MIN=8 timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2s_batch1_pairs8.log 2>&1
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2s_profile_layer_sweep.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2s_profile_weighted_subtree.log 2>&1
timeout 300 python tools/generic_profile.py lightcnn > gpurun_out/r2s_profile_lightcnn.log 2>&1
grep -v "^$" gpurun_out/r2s_pytest.log | tail -n 8 | cut -c1-300
tail -n 1 gpurun_out/r2s_batch1.log gpurun_out/r2s_batch1_bn256.log gpurun_out/r2s_batch1_pairs.log gpurun_out/r2s_batch1_pairs8.log
grep -A 12 "ms per call" gpurun_out/r2s_profile_layer_sweep.log | cut -c1-170
grep -A 9 "ms per call" gpurun_out/r2s_profile_weighted_subtree.log | cut -c1-170
grep -A 8 "ms per call" gpurun_out/r2s_profile_lightcnn.log | cut -c1-170
