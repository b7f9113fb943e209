#!/bin/bash
mkdir -p gpurun_out
KLAMPT_B200_OPTIONS="graph_max=16384" timeout 300 python scripts/gpu_latency2.py c2 2>&1 | head -8 > gpurun_out/latency_static_16k.log
cat gpurun_out/latency_static_16k.log
