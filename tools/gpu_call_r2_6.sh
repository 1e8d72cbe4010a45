#!/bin/bash
# GPU call 6: 12 epilogue warps + JOIN L2 prefetch; suite; ncu full captures
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2f_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2f_bf16_tests.log
XFRB_BENCH_LAUNCHES=gpurun_out/r2f_launches.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
XFRB_DBG=8 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2f_bench_noprefetch.json 2> gpurun_out/r2f_bench_noprefetch.err
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2f_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2f_pytest.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
B="python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3"
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)4, \(int\)3' -s 8 -c 1 -o gpurun_out/r2f_ncu_join -f $B > gpurun_out/r2f_ncu_join.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)4, \(int\)2' -s 20 -c 2 -o gpurun_out/r2f_ncu_mid -f $B > gpurun_out/r2f_ncu_mid.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)5, \(int\)1' -s 340 -c 3 -o gpurun_out/r2f_ncu_fwd -f $B > gpurun_out/r2f_ncu_fwd.log 2>&1
tail -4 gpurun_out/r2f_bf16_tests.log
grep -v "^$" gpurun_out/r2f_pytest.log | tail -12 | cut -c1-300
python tools/launch_roofline.py gpurun_out/r2f_launches.jsonl 2232 2 | head -16
for f in r2f_bench r2f_bench_noprefetch; do cut -c1-200 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
ls -la gpurun_out/*.ncu-rep | tail -4
