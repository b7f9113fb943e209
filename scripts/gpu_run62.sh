#!/bin/bash
mkdir -p gpurun_out
for w in c2 c3; do
  timeout 300 python bench.py --workload $w --extras 0 --cpu-seconds 1 --steps 8 > gpurun_out/bench_pref_$w.log 2>&1
  python - $w <<'PY'
import json,sys
w=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_pref_%s.log'%w) if l.startswith('{')][-1])
print(w,"value %.4g ms %.3f kernel %.3f e2e %.4g feas %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["e2e"]["value"],d["feasible_fraction"]))
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q --timeout=600 2>&1 | tail -3
