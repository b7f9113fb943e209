#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raycast.py -x -q --timeout=600 > gpurun_out/pytest_raycast.log 2>&1
tail -5 gpurun_out/pytest_raycast.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_raycast -s 6 -c 1 -o gpurun_out/prof_raycast_r02 -f python bench.py --extras 2 --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_raycast.log 2>&1
ncu -i gpurun_out/prof_raycast_r02.ncu-rep --page raw --csv > gpurun_out/prof_raycast_r02_raw.csv 2>/dev/null
ls -la gpurun_out/prof_raycast_r02*
