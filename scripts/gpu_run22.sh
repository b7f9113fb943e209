#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_golden.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_22.log 2>&1; tail -6 gpurun_out/pytest_22.log
V=$PWD/klampt_b200/_variants
for wl in c2 c3 c1; do
  B="timeout 300 python bench.py --extras 0 --cpu-seconds 1 --workload $wl"
  $B > gpurun_out/b22_${wl}_wide.log 2>&1
  KLAMPT_B200_LIB=$V/libklampt_b200_bps6.so $B > gpurun_out/b22_${wl}_wide_bps6.log 2>&1
done
timeout 300 python bench.py --extras 0 --cpu-seconds 1 --workload c4 --configs 200000 > gpurun_out/b22_c4_wide.log 2>&1
KLAMPT_B200_OPTIONS=wide=0 timeout 300 python bench.py --extras 0 --cpu-seconds 1 --workload c4 --configs 200000 > gpurun_out/b22_c4_binary.log 2>&1
for f in gpurun_out/b22_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); r=d.get("roofline") or {}
        print("  value %.4g ms %.3f e2e %.4g kernel_ms %.3f share %.3f feas %.4f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r.get("avg_launch_ms",0),r.get("kernel_share_of_step",0),d["feasible_fraction"]))
        break
else:
    print(open(sys.argv[1]).read()[-1500:])
PY
done
timeout 200 python scripts/gpu_stats.py c3 2>&1 | tail -2
