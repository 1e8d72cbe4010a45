#!/bin/bash
# GPU call 17: records of the extra workloads, Light-CNN kernel table, batch-1 latency with / without PDL + batch-1 launch list
mkdir -p gpurun_out
for w in layer_sweep weighted_subtree lightcnn; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2r_bench_$w.json 2> gpurun_out/r2r_bench_$w.err
done
timeout 300 python tools/generic_profile.py lightcnn > gpurun_out/r2r_profile_lightcnn.log 2>&1
timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2r_batch1_nopdl.log 2>&1
XFRB_PDL=1 timeout 200 python tools/batch1_profile.py 12 > gpurun_out/r2r_batch1_pdl.log 2>&1
XFRB_PDL=1 timeout 300 python -m pytest tests/test_bf16x2.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2r_pdl_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2r_pdl_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/r2r_ncu_launches_batch1.csv python tools/batch1_profile.py 4 > gpurun_out/r2r_ncu_batch1.log 2>&1
for w in layer_sweep weighted_subtree lightcnn; do cut -c1-230 gpurun_out/r2r_bench_$w.json; tail -n 2 gpurun_out/r2r_bench_$w.err | cut -c1-200; done
tail -n 20 gpurun_out/r2r_profile_lightcnn.log | cut -c1-160
tail -n 2 gpurun_out/r2r_batch1_nopdl.log gpurun_out/r2r_batch1_pdl.log gpurun_out/r2r_pdl_tests.log
wc -l gpurun_out/r2r_ncu_launches_batch1.csv
