#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py -x -q --timeout=600 2>&1 | tail -5
timeout 200 python scripts/gpu_latency2.py c2 2>&1 | head -3
