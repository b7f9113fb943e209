#!/bin/bash
mkdir -p gpurun_out
echo "zero_copy_max=0" > gpurun_out/latency_zc.log
KLAMPT_B200_OPTIONS="zero_copy_max=0" timeout 300 python scripts/gpu_latency2.py c2 2>&1 | head -4 >> gpurun_out/latency_zc.log
echo "zero_copy_max=64 (default)" >> gpurun_out/latency_zc.log
timeout 300 python scripts/gpu_latency2.py c2 2>&1 | head -4 >> gpurun_out/latency_zc.log
echo "zero_copy_max=1024" >> gpurun_out/latency_zc.log
KLAMPT_B200_OPTIONS="zero_copy_max=1024" timeout 300 python scripts/gpu_latency2.py c2 2>&1 | head -5 >> gpurun_out/latency_zc.log
cat gpurun_out/latency_zc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "small or graph" --timeout=600 2>&1 | tail -3
