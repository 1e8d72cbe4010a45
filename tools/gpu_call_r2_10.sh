#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/stream_mix > gpurun_out/r2j_stream_mix.jsonl 2>&1
cat gpurun_out/r2j_stream_mix.jsonl
timeout 600 python -m pytest tests -m gpu -q -rs -k "cubic or subtree_resnet101" > gpurun_out/r2j_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2j_pytest.log
tail -n 5 gpurun_out/r2j_pytest.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
B="python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3"
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)4, \(int\)3' -s 8 -c 1 -o gpurun_out/r2j_ncu_join -f $B > gpurun_out/r2j_ncu_join.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)4, \(int\)2' -s 20 -c 2 -o gpurun_out/r2j_ncu_mid -f $B > gpurun_out/r2j_ncu_mid.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)5, \(int\)1' -s 340 -c 3 -o gpurun_out/r2j_ncu_fwd -f $B > gpurun_out/r2j_ncu_fwd.log 2>&1
timeout 300 $NCU -k 'regex:join_kernel' -s 6 -c 1 -o gpurun_out/r2j_ncu_joink -f $B > gpurun_out/r2j_ncu_joink.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2j_ncu_launches.csv $B > gpurun_out/r2j_ncu_b.log 2>&1
ls -la gpurun_out/r2j*.ncu-rep
