#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raycast.py -x -q --timeout=600 > gpurun_out/pytest_raycast.log 2>&1
tail -15 gpurun_out/pytest_raycast.log
timeout 600 python bench.py --extras 2 --cpu-seconds 6 > gpurun_out/bench_rays.log 2>&1
tail -c 2500 gpurun_out/bench_rays.log
