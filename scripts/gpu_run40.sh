#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raycast.py -x -q --timeout=600 > gpurun_out/pytest_raycast.log 2>&1
tail -5 gpurun_out/pytest_raycast.log
timeout 300 python scripts/gpu_rays.py 0 > gpurun_out/rays_variants.log 2>&1
timeout 300 python scripts/gpu_rays.py 1 >> gpurun_out/rays_variants.log 2>&1
cat gpurun_out/rays_variants.log | tail -12
