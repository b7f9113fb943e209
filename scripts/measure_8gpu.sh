#!/bin/bash
# The round's 8-GPU record (gpurun --gpus 8): torchrun bench with extras, the one-handle multi-device script (only 1 and 8 devices to
# save box time), and the tests that need two GPUs.  usage: bash scripts/measure_8gpu.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --cpu-seconds 3 ) > gpurun_out/bench_8gpu_$TAG.log 2>&1
tail -c 400 gpurun_out/bench_8gpu_$TAG.log
( time timeout 600 python scripts/gpu_multi.py 8 1000000 18 ) > gpurun_out/multidevice_8gpu_$TAG.log 2>&1
tail -4 gpurun_out/multidevice_8gpu_$TAG.log
( time timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_sharding_nccl.py -q --timeout=600 ) > gpurun_out/pytest_2gpu_$TAG.log 2>&1
tail -3 gpurun_out/pytest_2gpu_$TAG.log
