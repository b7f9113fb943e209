"""option mesh_builder = 1: the merged environment mesh hierarchy from the GPU builder; finalize time, query time, identical answers"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from scripts.gpu_grid import timed
w = synth.world_c2()
t = time.time(); e0 = Engine(w); t0 = time.time() - t
t = time.time(); e1 = Engine(w, options={"mesh_builder": 1}); t1 = time.time() - t
print("C2 kb_finalize: host SAH %.2f s, GPU builder for the merged environment mesh %.2f s" % (t0, t1))
N = 1000000
Q = synth.sample_configs(w.robot, N, 2)
r0, r1 = e0.feasible_batch(Q), e1.feasible_batch(Q)
print("mismatches", int((r0 != r1).sum()), "feasible", r0.mean())
d0, d1 = e0.distance_batch(Q[:3000], upper_bound=0.3), e1.distance_batch(Q[:3000], upper_bound=0.3)
print("distance max |diff|", float(np.max(np.abs(d0 - d1))))
dQ = torch.from_numpy(Q).cuda(); dout = torch.empty(N, dtype=torch.uint8, device="cuda")
for name, eng in (("SAH", e0), ("LBVH", e1)):
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_option("collect_stats", 1); eng.reset_stats(); eng.feasible_batch_device(dQ, N, dout); torch.cuda.synchronize(); st = eng.stats(); eng.set_option("collect_stats", 0)
    print("  %-5s %.3f ms / 1M; per cfg: iter %.1f node %.0f elem %.1f" % (name, timed(eng, dQ, N, dout), st["node_iterations"] / N, st["node_tests"] / N, st["elem_tests"] / N))
