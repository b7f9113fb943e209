#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_raycast.py -x -q --timeout=600 2>&1 | tail -5
