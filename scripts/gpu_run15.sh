#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_golden.py -m gpu -q -x --timeout=120 > gpurun_out/pytest_15.log 2>&1; tail -6 gpurun_out/pytest_15.log
timeout 120 python scripts/gpu_latency2.py c2 2>&1 | tee gpurun_out/latency_c2_coop.log
KLAMPT_B200_OPTIONS=coop_max=0 timeout 120 python scripts/gpu_latency2.py c2 2>&1 | tee gpurun_out/latency_c2_nocoop.log
timeout 120 python scripts/gpu_latency.py 2>&1 | tail -3
