#!/bin/bash
# GPU call 4: un-gated suite + new parity tests, extra workloads, default bench with latency / library comparator, ncu full of the bf16x2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs > gpurun_out/r2d_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2d_pytest.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "wellcond or big_hooked" > gpurun_out/r2d_wellcond.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
for w in layer_sweep weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel<256, 4, 3" -s 8 -c 1 -o gpurun_out/r2d_ncu_join -f python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3 > gpurun_out/r2d_ncu_join.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel<256, 4, 2" -s 20 -c 2 -o gpurun_out/r2d_ncu_mid -f python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3 > gpurun_out/r2d_ncu_mid.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel<256, 5, 1" -s 40 -c 3 -o gpurun_out/r2d_ncu_fwd -f python bench.py --no-cpu-baseline --no-extras --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3 > gpurun_out/r2d_ncu_fwd.log 2>&1
tail -15 gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_wellcond.log | grep -v "^$" | tail -50
for f in default layer_sweep weighted_subtree lightcnn; do cut -c1-400 gpurun_out/r2d_bench_$f.json; tail -2 gpurun_out/r2d_bench_$f.err; done
