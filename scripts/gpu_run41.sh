#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --extras 0 --cpu-seconds 1 > gpurun_out/bench_stage_frac.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_stage_frac.log') if l.startswith('{')][-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"f32",d["e2e_f32_rows"]["value"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "staged or large or big or piece" --timeout=600 2>&1 | tail -3
