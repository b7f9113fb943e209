#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_stats.py c2 2>&1 | tee gpurun_out/stats_c2.log
timeout 200 python scripts/gpu_stats.py c3 2>&1 | tee gpurun_out/stats_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_traverse_wide -s 3 -c 1 -o gpurun_out/prof_traverse_wide_r02 -f python bench.py --steps 2 --warmup 3 --extras 0 --cpu-seconds 1 > gpurun_out/ncu_wide.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
