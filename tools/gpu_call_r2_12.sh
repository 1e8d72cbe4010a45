#!/bin/bash
# GPU call 12: JOIN operand staging by cp.async (A) against the previous JOIN epilogue (B = XFRB_PAIRA_JOIN_STAGE=0)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2l_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2l_bf16_tests.log
XFRB_BENCH_LAUNCHES=gpurun_out/r2l_launches_A.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_A.json 2> gpurun_out/r2l_bench_A.err
XFRB_LIB=$PWD/xfr_b200/libxfr_b200_nostage.so XFRB_BENCH_LAUNCHES=gpurun_out/r2l_launches_B.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_B.json 2> gpurun_out/r2l_bench_B.err
timeout 600 python -m pytest tests -m gpu -q -rs -k "parity or resnet50 or inpaint or subtree" > gpurun_out/r2l_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2l_pytest.log
tail -n 3 gpurun_out/r2l_bf16_tests.log; grep -v "^$" gpurun_out/r2l_pytest.log | tail -n 6 | cut -c1-250
for v in A B; do echo "VARIANT $v"; python tools/launch_roofline.py gpurun_out/r2l_launches_$v.jsonl 2232 2 2>/dev/null | grep dgrad_join; cut -c1-140 gpurun_out/r2l_bench_$v.json; tail -n 2 gpurun_out/r2l_bench_$v.err; done
