#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_sharding_nccl.py -q --timeout=600 ) > gpurun_out/pytest_2gpu_final.log 2>&1
tail -5 gpurun_out/pytest_2gpu_final.log
