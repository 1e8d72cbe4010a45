#!/bin/bash
# GPU call 31: ncu launch list + full captures of the end-of-round build (128-probe sweeps, as the earlier lists)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --batch 128 --chunk 128 --steps 1 --warmup 3"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2ah_ncu_launches_final.csv $B > gpurun_out/r2ah_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 400 $NCU -k 'regex:stem_bwd_b_kernel|contrast_kernel|join_kernel' -s 16 -c 6 -o gpurun_out/r2ah_ncu_elementwise -f $B > gpurun_out/r2ah_ncu_elementwise.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)256, \(int\)4, \(int\)3' -s 8 -c 1 -o gpurun_out/r2ah_ncu_join -f $B > gpurun_out/r2ah_ncu_join.log 2>&1
timeout 300 $NCU -k 'regex:conv_tc_kernel<\(int\)128, \(int\)4, \(int\)2' -s 20 -c 1 -o gpurun_out/r2ah_ncu_mid128 -f $B > gpurun_out/r2ah_ncu_mid128.log 2>&1
wc -l gpurun_out/r2ah_ncu_launches_final.csv; ls -la gpurun_out/r2ah_*.ncu-rep
