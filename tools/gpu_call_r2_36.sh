#!/bin/bash
# GPU call 36: compute-sanitizer memcheck over the kernels written in the second half of round 2 (hook chains / row walk / prior tables,
# many-block score + threshold kernels, arg-byte max-pool backward, stem quad kernel, join / contrast rewrites, Light-CNN stem + MFM fusion)
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20"
timeout 330 $CS python -m pytest tests/test_generic_sweeps.py -m gpu -q -x -k "chains and affineonly_with_prior" > gpurun_out/r2ao_memcheck_generic.log 2>&1; echo "rc $?" >> gpurun_out/r2ao_memcheck_generic.log
timeout 200 $CS python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ao_memcheck_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r2ao_memcheck_smoke.log
timeout 330 $CS python -m pytest tests/test_layerwise_subtree.py -m gpu -q -x -k "test_weighted_subtree_gpu or test_layer_sweep_gpu" > gpurun_out/r2ao_memcheck_subtree.log 2>&1; echo "rc $?" >> gpurun_out/r2ao_memcheck_subtree.log
for f in generic smoke subtree; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|rc |Invalid|smoke" gpurun_out/r2ao_memcheck_$f.log | tail -n 6 | cut -c1-200; done
