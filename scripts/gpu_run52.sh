#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_latency2.py c2 2>&1 | head -6 > gpurun_out/latency_static.log
cat gpurun_out/latency_static.log
timeout 300 python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_static.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_static.log') if l.startswith('{')][-1])
print("C2 value %.4g ms %.3f kernel %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["e2e"]["value"]))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py -x -q --timeout=600 2>&1 | tail -3
