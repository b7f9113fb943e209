#!/bin/bash
mkdir -p gpurun_out
for v in main lf4 lf8 lf24; do
  if [ $v = main ]; then unset KLAMPT_B200_LIB; else export KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_$v.so; fi
  timeout 300 python bench.py --extras 0 --cpu-seconds 0.5 > gpurun_out/bench_$v.log 2>&1
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_%s.log'%v) if l.startswith('{')][-1])
print(v,"C2 value %.4g ms %.3f kernel %.3f feas %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["feasible_fraction"]))
PY
done
