#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_golden.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_23.log 2>&1; tail -6 gpurun_out/pytest_23.log
timeout 300 python scripts/gpu_dist2.py 8 2>&1 | tee gpurun_out/dist2_wide.log
KLAMPT_B200_OPTIONS=wide=0 timeout 300 python scripts/gpu_dist2.py 8 2>&1 | tail -3 | tee gpurun_out/dist2_binary.log
timeout 300 python scripts/gpu_dist.py c2 2>&1 | tee gpurun_out/dist_c2_wide.log
V=$PWD/klampt_b200/_variants
for wl in c2 c1; do
  B="timeout 300 python bench.py --extras 0 --cpu-seconds 1 --workload $wl"
  $B > gpurun_out/b23_${wl}_wide.log 2>&1
  KLAMPT_B200_LIB=$V/libklampt_b200_wpf.so $B > gpurun_out/b23_${wl}_wide_pf.log 2>&1
done
for f in gpurun_out/b23_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); r=d.get("roofline") or {}
        print("  value %.4g ms %.3f e2e %.4g kernel_ms %.3f share %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r.get("avg_launch_ms",0),r.get("kernel_share_of_step",0)))
        break
else:
    print(open(sys.argv[1]).read()[-1500:])
PY
done
