#!/bin/bash
mkdir -p gpurun_out
for v in pipe5 pipe4; do
  if [ $v = main ]; then unset KLAMPT_B200_LIB; else export KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_$v.so; fi
  timeout 300 python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_$v.log 2>&1
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench_%s.log'%v) if l.startswith('{')][-1])
    print(v,"value %.4g ms %.3f kernel %.3f e2e %.4g feas %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["e2e"]["value"],d["feasible_fraction"]))
except Exception as ex:
    print(v,"failed",ex); print(open('gpurun_out/bench_%s.log'%v).read()[-1500:])
PY
  timeout 300 python bench.py --workload c3 --extras 0 --cpu-seconds 1 --steps 5 > gpurun_out/bench_c3_$v.log 2>&1
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench_c3_%s.log'%v) if l.startswith('{')][-1])
    print(v,"C3 value %.4g ms %.3f kernel %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"]))
except Exception as ex:
    print(v,"failed",ex)
PY
done
export KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_pipe5.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 2>&1 | tail -3
