"""edge-checking throughput and the underlying configuration-check rate (device-resident)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = synth.world_c2()
eng = Engine(w)
A, B = synth.sample_edges(w.robot, lambda Q: eng.feasible_batch(Q), N, 4)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
out = torch.empty(N, dtype=torch.uint8, device="cuda"); nch = torch.empty(N, dtype=torch.int32, device="cuda")
s = torch.cuda.Stream(); eng.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    eng.edges_visible_batch_device(dA, dB, N, 0.01, out, nch); torch.cuda.synchronize()
    eng.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3): eng.edges_visible_batch_device(dA, dB, N, 0.01, out, nch)
    e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
st = eng.stats()
print("edges %d: %.2f ms -> %.3e edges/s; visible %.3f; config checks per call %.3e (level-synchronous) -> %.3e cfg/s; sequential-checker checks/edge %.1f; launches/call %d"
      % (N, ms, N / ms * 1e3, out.float().mean().item(), st["edge_config_checks"] / 3, st["edge_config_checks"] / 3 / ms * 1e3, nch.float().mean().item(), st["kernel_launches"] / 3))
