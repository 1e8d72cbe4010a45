#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call_r2.sh'
# Each step runs under its own timeout (a kernel instantiation that never ran must not be able to hang the box); outputs land
# in gpurun_out/.
mkdir -p gpurun_out
# 1. the verified suite (53 tests at the end of round 1)
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r2a_pytest.log
# 2. the gated tests: layer sweep, job functions vs the reference, the two-pass forward plan (conv_tc_kernel<BN, 2, FWD_DUAL>)
XFRB_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests -m gpu -q -k "layer_sweep or jobs_vs_reference or two_pass_forward" \
    > gpurun_out/r2a_unverified.log 2>&1; echo "rc $?" >> gpurun_out/r2a_unverified.log
# 3. bench: default plan, two-pass forward plan, 5 CTAs per SM in join_kernel
timeout 240 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
timeout 240 python bench.py --no-cpu-baseline --gemm tf32x2f > gpurun_out/r2a_bench_tf32x2f.json 2> gpurun_out/r2a_bench_tf32x2f.err
XFRB_JOIN=5 timeout 240 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_join5.json 2> gpurun_out/r2a_bench_join5.err
# 4. the batched scoring path
timeout 120 python tools/scoring_bench.py > gpurun_out/r2a_scoring.log 2>&1
tail -3 gpurun_out/r2a_pytest.log gpurun_out/r2a_unverified.log
for f in default tf32x2f join5; do cut -c1-160 gpurun_out/r2a_bench_$f.json; done
