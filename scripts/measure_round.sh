#!/bin/bash
# The round's record on one B200 (gpurun): whole GPU test suite, default bench (headline + extras), reference arm, launch list,
# ncu --set full of the two dominant kernels.  usage: bash scripts/measure_round.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=900 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_$TAG.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_reference_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --extras 0 --cpu-seconds 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_traverse -s 3 -c 1 -o gpurun_out/prof_traverse_$TAG -f python bench.py --steps 2 --warmup 3 --extras 0 --cpu-seconds 1 > gpurun_out/ncu_traverse_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_raycast -s 6 -c 1 -o gpurun_out/prof_raycast_$TAG -f python bench.py --extras 2 --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_raycast_$TAG.log 2>&1
ncu -i gpurun_out/prof_raycast_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raycast_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_traverse_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_traverse_${TAG}_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -c 600 gpurun_out/bench_$TAG.log
