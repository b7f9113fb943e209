#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_golden.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_17.log 2>&1; tail -6 gpurun_out/pytest_17.log
timeout 120 python scripts/gpu_latency.py 2>&1 | tee gpurun_out/latency_edges_flat.log
KLAMPT_B200_OPTIONS=edge_flat_max=0 timeout 120 python scripts/gpu_latency.py 2>&1 | tail -2 | tee gpurun_out/latency_edges_level.log
