#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:kb_distance -s 1 -c 1 -o gpurun_out/prof_distance_r02_v2 -f python scripts/gpu_dist2.py 8 50000 > gpurun_out/ncu_dist2.log 2>&1
ls -la gpurun_out/*.ncu-rep
