#!/bin/bash
# C2 headline at 2 and 4 GPUs on one 4-GPU box (gpurun --gpus 4): fills the 1 / 2 / 4 / 8 scaling table
mkdir -p gpurun_out
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --extras 0 --cpu-seconds 1 > gpurun_out/bench_${n}gpu.log 2>&1
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_%sgpu.log'%n) if l.startswith('{')][-1])
print(n,"GPUs: value %.4g ms %.3f e2e %.4g f32 %.4g strong %.4g"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e_f32_rows"]["value"],d["strong"]["value"]))
PY
done
