#!/bin/bash
# 8-GPU records of the BASELINE workloads with the final build (one process per GPU, NCCL); Light-CNN (configs[4]) on 4 of them
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8"
timeout 300 $RUN --steps 5 --warmup 3 > gpurun_out/r2aj_bench8_contrastive.json 2> gpurun_out/r2aj_bench8_contrastive.err
timeout 300 $RUN --workload weighted_subtree --steps 3 --warmup 6 > gpurun_out/r2aj_bench8_weighted_subtree.json 2> gpurun_out/r2aj_bench8_weighted_subtree.err
timeout 300 $RUN --workload layer_sweep --steps 3 --warmup 6 > gpurun_out/r2aj_bench8_layer_sweep.json 2> gpurun_out/r2aj_bench8_layer_sweep.err
RUN4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4"
timeout 300 $RUN4 --workload lightcnn --steps 3 --warmup 3 > gpurun_out/r2aj_bench4_lightcnn.json 2> gpurun_out/r2aj_bench4_lightcnn.err
for w in bench8_contrastive bench8_weighted_subtree bench8_layer_sweep bench4_lightcnn; do grep -h '"metric"' gpurun_out/r2aj_$w.json | cut -c1-200; done
