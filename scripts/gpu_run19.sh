#!/bin/bash
# 8 GPUs of one box: the driver's launch of bench.py (one process per GPU, NCCL), then one process / one handle over all GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
N=$(nvidia-smi -L | wc -l)
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/bench_${N}gpu.log 2>&1
tail -c 400 gpurun_out/bench_${N}gpu.log
timeout 600 python scripts/gpu_multi.py $N 2>&1 | tee gpurun_out/multi_${N}gpu.log
