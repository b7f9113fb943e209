#!/bin/bash
mkdir -p gpurun_out
for v in head main; do
  if [ $v = main ]; then unset KLAMPT_B200_LIB; else export KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_$v.so; fi
  echo "== $v" | tee -a gpurun_out/latency_mid.log
  timeout 300 python scripts/gpu_latency2.py c2 2>&1 | tail -7 | tee -a gpurun_out/latency_mid.log
  timeout 300 python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_mid_$v.log 2>&1
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_mid_%s.log'%v) if l.startswith('{')][-1])
print(v,"C2 value %.4g ms %.3f kernel %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["e2e"]["value"]))
PY
done
