"""Strong-scaling shard + gather on hardware (BASELINE config 3): two processes, one GPU each, NCCL.  One batch is split
with shard_range, every rank checks its block through the C ABI, the result bytes / bitmasks are all-gathered device-resident, and
the gathered answer must equal the one-GPU answer of the whole batch bit for bit.  Needs two GPUs (gpurun --gpus 2); the CPU-side
logic of the same code runs on gloo in tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["KB_ROOT"])
from klampt_b200 import synth
from klampt_b200.engine import Engine
from klampt_b200.shard import ShardedRunner, shard_range, interleaved_indices, gather_results
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
w = synth.world_c3()
eng = Engine(w, device=rank)
Q = synth.sample_configs(w.robot, 200_001, 31)              # ragged on purpose: not a multiple of the world size
run = ShardedRunner(feasible_fn=eng.feasible_batch, visible_fn=lambda A, B: eng.edges_visible_batch(A, B, eps=0.05, return_nchecks=False), device="cuda")
got = run.feasible_batch(Q)
# packed bitmasks, gathered as bytes
lo, hi = shard_range(len(Q), rank, world)
bits = eng.feasible_batch_bits(Q[lo:hi])
per = -(-len(Q) // world)
allbits = gather_results(np.pad(bits, (0, (per + 7) // 8 - len(bits))), ((per + 7) // 8) * world, device="cuda")
unp = np.concatenate([np.unpackbits(allbits[r * ((per + 7) // 8):(r + 1) * ((per + 7) // 8)], bitorder="little")[:shard_range(len(Q), r, world)[1] - shard_range(len(Q), r, world)[0]] for r in range(world)])
A, B = Q[:30_000], Q[30_000:60_000]
vis = run.visible_batch(A, B)
d = gather_results(eng.distance_batch(Q[lo:hi][:20_000 // world], upper_bound=0.3, include_self=True), (20_000 // world) * world, device="cuda")
if rank == 0:
    whole = eng.feasible_batch(Q)
    assert np.array_equal(got, whole), "gathered feasibility differs from the one-GPU answer"
    assert np.array_equal(unp, whole), "gathered bitmask differs from the one-GPU answer"
    assert np.array_equal(vis, eng.edges_visible_batch(A, B, eps=0.05, return_nchecks=False)), "gathered visibility differs"
    ref = np.concatenate([eng.distance_batch(Q[shard_range(len(Q), r, world)[0]:][:20_000 // world], upper_bound=0.3, include_self=True) for r in range(world)])
    assert np.array_equal(d, ref), "gathered distances differ"
    print("NCCL_SHARD_OK feasible=%.3f visible=%.3f" % (whole.mean(), vis.mean()))
dist.destroy_process_group()
'''


def test_two_gpu_shard_and_gather_equals_one_gpu(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_SHARD_OK" in r.stdout
