#!/bin/bash
# GPU call 8: A/B of the bf16x2 epilogue knobs. A = product build (8 epilogue warps, paired slabs, JOIN load-ahead); B = A + TMEM prefetch;
# C = 12 epilogue warps, paired; D = A with unpaired slabs
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2h_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2h_bf16_tests.log
for v in A B C D; do
  L=""; [ $v != A ] && L=$PWD/xfr_b200/libxfr_b200_$v.so
  XFRB_LIB=$L XFRB_BENCH_LAUNCHES=gpurun_out/r2h_launches_$v.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_$v.json 2> gpurun_out/r2h_bench_$v.err
done
XFRB_LIB=$PWD/xfr_b200/libxfr_b200_B.so timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2h_bf16_tests_B.log 2>&1; echo "rc $?" >> gpurun_out/r2h_bf16_tests_B.log
tail -n 3 gpurun_out/r2h_bf16_tests.log gpurun_out/r2h_bf16_tests_B.log
for v in A B C D; do echo "VARIANT $v"; python tools/launch_roofline.py gpurun_out/r2h_launches_$v.jsonl 2232 2 2>/dev/null | sed -n 2,8p; cut -c1-140 gpurun_out/r2h_bench_$v.json; tail -n 2 gpurun_out/r2h_bench_$v.err; done
