#!/bin/bash
# round 2, GPU call 1: whole GPU test suite under the two-sided band rule, default bench (headline + extras), reference arm,
# launch list, and the small-shared-memory variant of the traversal kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -q -x --timeout=1500 ) > gpurun_out/pytest_gpu.log 2>&1
( time python bench.py ) > gpurun_out/bench_default.log 2>&1
( time python bench.py --impl reference ) > gpurun_out/bench_reference.log 2>&1
KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_s320.so python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_s320.log 2>&1
python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_base2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_a.csv python bench.py --steps 2 --warmup 1 --extras 0 --cpu-seconds 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
