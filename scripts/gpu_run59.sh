#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_raycast.py -x -q -k "dynamic or cloud or raycast or ray or camera or adapters" --timeout=600 2>&1 | tail -8
