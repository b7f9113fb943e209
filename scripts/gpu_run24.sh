#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_golden.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_24.log 2>&1; tail -6 gpurun_out/pytest_24.log
timeout 300 python scripts/gpu_dist2.py 8 2>&1 | tee gpurun_out/dist2_wide.log
timeout 300 python scripts/gpu_dist.py c2 2>&1 | tee gpurun_out/dist_c2_wide.log
KLAMPT_B200_OPTIONS=wide=0 timeout 300 python scripts/gpu_dist.py c2 2>&1 | tee gpurun_out/dist_c2_binary.log
