#!/bin/bash
# GPU call 38: compute-sanitizer racecheck (shared-memory hazards) over the new shared-memory kernels
mkdir -p gpurun_out
CS="compute-sanitizer --tool racecheck --error-exitcode 86 --print-limit 20"
timeout 400 $CS python -m pytest tests/test_layerwise_subtree.py -m gpu -q -x -k "test_layer_sweep_gpu" > gpurun_out/r2aq_racecheck_sweep.log 2>&1; echo "rc $?" >> gpurun_out/r2aq_racecheck_sweep.log
timeout 300 $CS python -m pytest tests/test_gpu_lightcnn.py -m gpu -q -x -k "test_vs_reference and tf32x3 and affineonly_with_prior" > gpurun_out/r2aq_racecheck_lightcnn.log 2>&1; echo "rc $?" >> gpurun_out/r2aq_racecheck_lightcnn.log
for f in sweep lightcnn; do echo "== $f"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|rc |hazard" gpurun_out/r2aq_racecheck_$f.log | tail -n 6 | cut -c1-220; done
