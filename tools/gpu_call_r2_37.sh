#!/bin/bash
# GPU call 37: compute-sanitizer memcheck, Light-CNN (stem conv rewrite, MFM fused into the Split firing) and the VGGFace2 ResNet-50 (pool pad 0)
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20"
timeout 330 $CS python -m pytest tests/test_gpu_lightcnn.py -m gpu -q -x -k "test_vs_reference and tf32x3 and affineonly" > gpurun_out/r2ap_memcheck_lightcnn.log 2>&1; echo "rc $?" >> gpurun_out/r2ap_memcheck_lightcnn.log
timeout 330 $CS python -m pytest tests/test_resnet50_128.py -m gpu -q -x > gpurun_out/r2ap_memcheck_r50.log 2>&1; echo "rc $?" >> gpurun_out/r2ap_memcheck_r50.log
timeout 200 $CS python -m pytest tests/test_stream.py -m gpu -q -x > gpurun_out/r2ap_memcheck_stream.log 2>&1; echo "rc $?" >> gpurun_out/r2ap_memcheck_stream.log
for f in lightcnn r50 stream; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|skipped|rc |Invalid" gpurun_out/r2ap_memcheck_$f.log | tail -n 5 | cut -c1-200; done
