#!/bin/bash
# round 2, GPU call 5: how often comparable inner pairs descend both trees at once (KB_BOTH_RATIO), C2 and C3
mkdir -p gpurun_out
V=$PWD/klampt_b200/_variants
for wl in c2 c3; do
  B="python bench.py --extras 0 --cpu-seconds 1 --workload $wl"
  $B > gpurun_out/b5_${wl}_br4.log 2>&1
  for r in 16 64 inf; do KLAMPT_B200_LIB=$V/libklampt_b200_br$r.so $B > gpurun_out/b5_${wl}_br$r.log 2>&1; done
done
for f in gpurun_out/b5_*.log; do echo $f; python - "$f" <<'PY'
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
