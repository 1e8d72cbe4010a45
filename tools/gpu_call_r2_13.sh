#!/bin/bash
# GPU call 13: launch list of the Light-CNN workload
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2n_ncu_launches_lightcnn.csv python bench.py --workload lightcnn --batch 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2n_ncu_lc.log 2>&1
tail -n 2 gpurun_out/r2n_ncu_lc.log | cut -c1-300
wc -l gpurun_out/r2n_ncu_launches_lightcnn.csv
