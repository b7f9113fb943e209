#!/bin/bash
# 8 GPUs, headline only: effect of the 8-piece staged upload on the end-to-end figure
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --extras 0 --cpu-seconds 1 > gpurun_out/bench_8gpu_staged8.log 2>&1
tail -c 1500 gpurun_out/bench_8gpu_staged8.log
