#!/bin/bash
# 4-GPU record of BASELINE configs[4]: Light-CNN-29v2 EBP, batch 512 per GPU
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4"
timeout 300 $RUN --workload lightcnn --steps 3 --warmup 2 > gpurun_out/r2m_bench4_lightcnn.json 2> gpurun_out/r2m_bench4_lightcnn.err
grep -h '"metric"' gpurun_out/r2m_bench4_lightcnn.json | cut -c1-260; tail -n 2 gpurun_out/r2m_bench4_lightcnn.err | cut -c1-200
