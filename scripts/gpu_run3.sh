#!/bin/bash
# round 2, GPU call 3: ncu --set full with source-level counters of the boolean traversal kernel (C2, 1 M configurations per launch)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:kb_traverse -s 3 -c 1 -o gpurun_out/prof_traverse_r02_base -f \
    python bench.py --steps 2 --warmup 3 --extras 0 --cpu-seconds 1 > gpurun_out/ncu_base.log 2>&1
ls -la gpurun_out/*.ncu-rep
