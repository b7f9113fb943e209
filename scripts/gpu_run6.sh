#!/bin/bash
# round 2, GPU call 6: distance kernel v2 -- parity tests that use distances, then C5 with cloud leaves of 8 / 16 / 32 points
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_adapters.py -m gpu -q --timeout=900 > gpurun_out/pytest_dist2.log 2>&1; tail -15 gpurun_out/pytest_dist2.log
for leaf in 8 16 32; do python scripts/gpu_dist2.py $leaf 2>&1 | tee gpurun_out/dist2_leaf$leaf.log; done
