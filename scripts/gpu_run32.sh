#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raycast.py -x -q --timeout=600 > gpurun_out/pytest_raycast.log 2>&1
tail -30 gpurun_out/pytest_raycast.log
