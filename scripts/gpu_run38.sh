#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_rays.py > gpurun_out/rays_variants.log 2>&1
cat gpurun_out/rays_variants.log | tail -15
