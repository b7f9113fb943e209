#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py -m gpu -q --timeout=900 > gpurun_out/pytest_14.log 2>&1; tail -6 gpurun_out/pytest_14.log
python scripts/gpu_latency2.py c2 2>&1 | tee gpurun_out/latency_c2_before.log
python scripts/gpu_latency2.py c1 2>&1 | tee gpurun_out/latency_c1_before.log
