#!/bin/bash
# GPU call 5: the whole suite on the bf16x2 default + ncu full captures of the three bf16x2 kernel families
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2e_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2e_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
B="python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3"
timeout 300 $NCU -k regex:"conv_tc_kernel<256, 4, 3" -s 8 -c 1 -o gpurun_out/r2e_ncu_join -f $B > gpurun_out/r2e_ncu_join.log 2>&1
timeout 300 $NCU -k regex:"conv_tc_kernel<256, 4, 2" -s 20 -c 2 -o gpurun_out/r2e_ncu_mid -f $B > gpurun_out/r2e_ncu_mid.log 2>&1
timeout 300 $NCU -k regex:"conv_tc_kernel<256, 5, 1" -s 340 -c 3 -o gpurun_out/r2e_ncu_fwd -f $B > gpurun_out/r2e_ncu_fwd.log 2>&1
grep -v "^$" gpurun_out/r2e_pytest.log | tail -40 | cut -c1-300
cat gpurun_out/r2e_smoke.log | tail -3
cut -c1-300 gpurun_out/r2e_bench_default.json; tail -3 gpurun_out/r2e_bench_default.err
ls -la gpurun_out/*.ncu-rep | tail -4; tail -3 gpurun_out/r2e_ncu_join.log | cut -c1-200
