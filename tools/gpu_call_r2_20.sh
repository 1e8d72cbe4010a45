#!/bin/bash
# GPU call 20: JOIN with 16 epilogue warps (A) against the default 12 (B); ncu of the chained hook kernel
mkdir -p gpurun_out
XFRB_LIB=$PWD/xfr_b200/libxfr_b200_join16.so timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2u_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2u_bf16_tests.log
for v in A B A B; do
  if [ $v = A ]; then export XFRB_LIB=$PWD/xfr_b200/libxfr_b200_join16.so; else unset XFRB_LIB; fi
  XFRB_BENCH_LAUNCHES=gpurun_out/r2u_launches_$v.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2u_bench_$v.json 2> gpurun_out/r2u_bench_$v.err
  echo "VARIANT $v"; python tools/launch_roofline.py gpurun_out/r2u_launches_$v.jsonl 2232 2 2>/dev/null | grep dgrad_join | head -3; cut -c1-120 gpurun_out/r2u_bench_$v.json
done
unset XFRB_LIB
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled --graph-profiling node -k regex:hook_kernel -s 700 -c 4 -o gpurun_out/r2u_ncu_hook -f python tools/generic_profile.py weighted_subtree > gpurun_out/r2u_ncu_hook.log 2>&1
tail -n 2 gpurun_out/r2u_bf16_tests.log; ls -la gpurun_out/r2u_ncu_hook.ncu-rep
