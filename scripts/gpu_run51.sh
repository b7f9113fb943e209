#!/bin/bash
mkdir -p gpurun_out
for v in main c3b4; do
  if [ $v = main ]; then unset KLAMPT_B200_LIB; else export KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_$v.so; fi
  timeout 300 python bench.py --workload c3 --extras 0 --cpu-seconds 1 --steps 5 > gpurun_out/bench_c3_$v.log 2>&1
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench_c3_%s.log'%v) if l.startswith('{')][-1])
    print(v,"C3 value %.4g ms %.3f kernel %.3f feas %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["feasible_fraction"]))
except Exception as ex:
    print(v,"failed",ex); print(open('gpurun_out/bench_c3_%s.log'%v).read()[-800:])
PY
done
