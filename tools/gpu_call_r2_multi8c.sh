#!/bin/bash
# 8-GPU record of the weighted-subtree workload with the corrected job sharding (bench.py: one global job list, T jobs per rank)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8"
timeout 300 $RUN --workload weighted_subtree --steps 3 --warmup 3 > gpurun_out/r2y_bench8_weighted_subtree.json 2> gpurun_out/r2y_bench8_weighted_subtree.err
grep -h '"metric"' gpurun_out/r2y_bench8_weighted_subtree.json | cut -c1-300; tail -n 2 gpurun_out/r2y_bench8_weighted_subtree.err | cut -c1-200
