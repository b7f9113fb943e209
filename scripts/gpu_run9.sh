#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_adapters.py -m gpu -q --timeout=900 > gpurun_out/pytest_cp.log 2>&1; tail -30 gpurun_out/pytest_cp.log
python scripts/gpu_dist2.py 8 2>&1 | tee gpurun_out/dist2_leaf8_b.log
