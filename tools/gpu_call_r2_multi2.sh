#!/bin/bash
# 2-GPU run of the default bench as the driver launches it: stdout must be the one JSON line
mkdir -p gpurun_out
echo "NCCL_DEBUG=$NCCL_DEBUG"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ak_bench2.json 2> gpurun_out/r2ak_bench2.err
wc -l gpurun_out/r2ak_bench2.json; cut -c1-160 gpurun_out/r2ak_bench2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2ak_bench2_ref.json 2> gpurun_out/r2ak_bench2_ref.err
wc -l gpurun_out/r2ak_bench2_ref.json; cut -c1-160 gpurun_out/r2ak_bench2_ref.json
