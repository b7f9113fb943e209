#!/bin/bash
# round 2, GPU call 2: coherence sort + CTA-shared chunks: quick parity subset, then C2 bench variants
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x --timeout=900 > gpurun_out/pytest_sort.log 2>&1; tail -2 gpurun_out/pytest_sort.log
B="python bench.py --extras 0 --cpu-seconds 1"
V=$PWD/klampt_b200/_variants
$B > gpurun_out/b_sort_w4.log 2>&1
KLAMPT_B200_OPTIONS=sort_configs=0 $B > gpurun_out/b_nosort_w4.log 2>&1
KLAMPT_B200_LIB=$V/libklampt_b200_c8.so $B > gpurun_out/b_sort_w4_c8.log 2>&1
KLAMPT_B200_LIB=$V/libklampt_b200_w10.so $B > gpurun_out/b_sort_w10.log 2>&1
KLAMPT_B200_LIB=$V/libklampt_b200_w20.so $B > gpurun_out/b_sort_w20.log 2>&1
KLAMPT_B200_LIB=$V/libklampt_b200_w20.so KLAMPT_B200_OPTIONS=sort_configs=0 $B > gpurun_out/b_nosort_w20.log 2>&1
for f in gpurun_out/b_*.log; do echo $f; python - "$f" <<'PY'
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
