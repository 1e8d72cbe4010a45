#!/bin/bash
# GPU call 2 of round 2: first run of the bf16x2 plan (pair tensors, tcgen05 kind::f16)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bf16x2.py -m gpu -q -x -s > gpurun_out/r2b_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2b_bf16_tests.log
timeout 200 python bench.py --no-cpu-baseline --gemm bf16x2 > gpurun_out/r2b_bench_bf16x2.json 2> gpurun_out/r2b_bench_bf16x2.err
XFRB_CTA2=0 timeout 200 python bench.py --no-cpu-baseline --gemm bf16x2 > gpurun_out/r2b_bench_bf16x2_nopair.json 2> gpurun_out/r2b_bench_bf16x2_nopair.err
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2b_pytest.log
tail -60 gpurun_out/r2b_bf16_tests.log
tail -5 gpurun_out/r2b_pytest.log
for f in bf16x2 bf16x2_nopair; do cut -c1-200 gpurun_out/r2b_bench_$f.json; tail -3 gpurun_out/r2b_bench_$f.err; done
