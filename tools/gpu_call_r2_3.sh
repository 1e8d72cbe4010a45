#!/bin/bash
mkdir -p gpurun_out
XFRB_BENCH_LAUNCHES=gpurun_out/r2c_launches_bf16x2.jsonl timeout 200 python bench.py --no-cpu-baseline --gemm bf16x2 > gpurun_out/r2c_bench_bf16x2.json 2> gpurun_out/r2c_bench_bf16x2.err
XFRB_BENCH_LAUNCHES=gpurun_out/r2c_launches_tf32x3.jsonl timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_tf32x3.json 2> gpurun_out/r2c_bench_tf32x3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2c_ncu_launches_bf16x2.csv python bench.py --no-cpu-baseline --gemm bf16x2 --batch 128 --chunk 128 --steps 1 --warmup 3 > gpurun_out/r2c_ncu_b.log 2>&1
python tools/launch_roofline.py gpurun_out/r2c_launches_bf16x2.jsonl 2232 2 | head -50
