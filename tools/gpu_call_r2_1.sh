#!/bin/bash
# GPU call 1 of round 2: the gated tests of round 1, the measured tensor peaks, the opt-in plans, the library comparator.
mkdir -p gpurun_out
timeout 120 tools/mma_peak > gpurun_out/r2a_mma_peak.jsonl 2>&1
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2a_pytest.log
XFRB_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests -m gpu -q -k "layer_sweep or jobs_vs_reference or two_pass_forward" \
    > gpurun_out/r2a_unverified.log 2>&1; echo "rc $?" >> gpurun_out/r2a_unverified.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
timeout 200 python bench.py --no-cpu-baseline --gemm tf32x2f > gpurun_out/r2a_bench_tf32x2f.json 2> gpurun_out/r2a_bench_tf32x2f.err
XFRB_JOIN=5 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_join5.json 2> gpurun_out/r2a_bench_join5.err
timeout 200 python tools/torch_eager_baseline.py 16 64 > gpurun_out/r2a_torch_eager.jsonl 2> gpurun_out/r2a_torch_eager.err
timeout 200 python tools/torch_eager_baseline.py 32 64 >> gpurun_out/r2a_torch_eager.jsonl 2>> gpurun_out/r2a_torch_eager.err
cat gpurun_out/r2a_mma_peak.jsonl
tail -3 gpurun_out/r2a_pytest.log gpurun_out/r2a_unverified.log
for f in default tf32x2f join5; do cut -c1-200 gpurun_out/r2a_bench_$f.json; done
cat gpurun_out/r2a_torch_eager.jsonl; tail -2 gpurun_out/r2a_torch_eager.err
