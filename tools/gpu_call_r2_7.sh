#!/bin/bash
# GPU call 7: no per-tile barrier in the bf16x2 epilogues; profiling switches (XFRB_DBG: 1 no epilogue loads, 2 no stores, 4 no hook math)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_bf16x2.py -m gpu -q -x > gpurun_out/r2g_bf16_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2g_bf16_tests.log
for d in 0 1 2 3 4 7; do
XFRB_DBG=$d XFRB_BENCH_LAUNCHES=gpurun_out/r2g_launches_dbg$d.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_dbg$d.json 2> gpurun_out/r2g_bench_dbg$d.err
done
timeout 600 python -m pytest tests -m gpu -q -rs -k "subtree_resnet101 or graph or api or jobs_vs or eps" > gpurun_out/r2g_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_bf16_tests.log
grep -v "^$" gpurun_out/r2g_pytest.log | tail -12 | cut -c1-300
for d in 0 1 2 3 4 7; do echo "DBG $d"; python tools/launch_roofline.py gpurun_out/r2g_launches_dbg$d.jsonl 2232 2 2>/dev/null | sed -n 2,8p; cut -c1-140 gpurun_out/r2g_bench_dbg$d.json; done
