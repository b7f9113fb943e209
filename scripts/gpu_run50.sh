#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapters.py tests/test_gpu_raycast.py -x -q --timeout=600 > gpurun_out/pytest_adapters.log 2>&1
tail -12 gpurun_out/pytest_adapters.log
