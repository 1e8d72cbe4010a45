#!/bin/bash
# 8-GPU records of the BASELINE workloads (one process per GPU, NCCL): configs[1] contrastive, configs[2] layer sweep, configs[3] weighted subtree
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8"
timeout 300 $RUN --steps 5 --warmup 3 > gpurun_out/r2m_bench8_contrastive.json 2> gpurun_out/r2m_bench8_contrastive.err
timeout 400 $RUN --workload weighted_subtree --steps 2 --warmup 1 > gpurun_out/r2m_bench8_weighted_subtree.json 2> gpurun_out/r2m_bench8_weighted_subtree.err
timeout 400 $RUN --workload layer_sweep --steps 2 --warmup 1 > gpurun_out/r2m_bench8_layer_sweep.json 2> gpurun_out/r2m_bench8_layer_sweep.err
for w in contrastive weighted_subtree layer_sweep; do grep -h '"metric"' gpurun_out/r2m_bench8_$w.json | cut -c1-260; tail -n 2 gpurun_out/r2m_bench8_$w.err | cut -c1-200; done
