#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_sharding_nccl.py -m gpu -q --timeout=1200 > gpurun_out/pytest_2gpu.log 2>&1; tail -25 gpurun_out/pytest_2gpu.log
python scripts/gpu_multi.py 2 2>&1 | tee gpurun_out/multi_2gpu.log
