#!/bin/bash
# round 2, GPU call 4: register prefetch of the next configuration's transforms (+ grab size variants), C2 and C3
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=900 > gpurun_out/pytest_pf.log 2>&1; tail -2 gpurun_out/pytest_pf.log
V=$PWD/klampt_b200/_variants
for wl in c2 c3; do
  B="python bench.py --extras 0 --cpu-seconds 1 --workload $wl"
  $B > gpurun_out/b4_${wl}_pf_g8.log 2>&1
  KLAMPT_B200_LIB=$V/libklampt_b200_g16.so $B > gpurun_out/b4_${wl}_pf_g16.log 2>&1
  KLAMPT_B200_LIB=$V/libklampt_b200_g32.so $B > gpurun_out/b4_${wl}_pf_g32.log 2>&1
done
for f in gpurun_out/b4_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); r=d.get("roofline") or {}
        print("  value %.4g ms %.3f e2e %.4g kernel_ms %.3f share %.3f launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r.get("avg_launch_ms",0),r.get("kernel_share_of_step",0),d["gpu_launches"]))
        break
else:
    print(open(sys.argv[1]).read()[-1500:])
PY
done
