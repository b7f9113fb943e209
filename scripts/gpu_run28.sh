#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x --timeout=600 -k "one_million or multi_chunk or fp32_conf or feasible_c1" > gpurun_out/pytest_28.log 2>&1; tail -4 gpurun_out/pytest_28.log
for wl in c2 c3; do timeout 300 python bench.py --extras 0 --cpu-seconds 1 --workload $wl > gpurun_out/b28_$wl.log 2>&1; done
for f in gpurun_out/b28_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); r=d.get("roofline") or {}
        print("  value %.4g ms %.3f e2e %.4g e2e_f32 %.4g kernel_ms %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],(d.get("e2e_f32_rows") or {}).get("value",0),r.get("avg_launch_ms",0)))
        break
else:
    print(open(sys.argv[1]).read()[-1500:])
PY
done
