#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_latency.py 2>&1 | tail -6 > gpurun_out/latency_edges_small.log
cat gpurun_out/latency_edges_small.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -q -x -k "edge or golden" --timeout=600 2>&1 | tail -3
