#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
