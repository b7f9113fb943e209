"""Cost of feasible vs infeasible configurations: time and statistics of the boolean query on each subset of a C2 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from scripts.gpu_grid import timed

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = {"c2": synth.world_c2, "c3": synth.world_c3}[name]()
eng = Engine(w)
N = 2000000
Q = synth.sample_configs(w.robot, N, 2)
res = eng.feasible_batch(Q).astype(bool)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
for label, sel in (("feasible", res), ("infeasible", ~res), ("all", np.ones(N, bool))):
    Qs = np.ascontiguousarray(Q[sel][:800000]); n = len(Qs)
    dQ = torch.from_numpy(Qs).cuda(); dout = torch.empty(n, dtype=torch.uint8, device="cuda")
    eng.set_option("collect_stats", 1); eng.reset_stats()
    eng.feasible_batch_device(dQ, n, dout); torch.cuda.synchronize()
    st = eng.stats(); eng.set_option("collect_stats", 0)
    ms = timed(eng, dQ, n, dout)
    print("%s %-10s n=%d: %.3f ms = %.2f us/1k cfg; per cfg: iter %.1f node %.1f elem %.1f recheck %.2f" % (
        name, label, n, ms, ms * 1e3 / (n / 1e3), st["node_iterations"] / n, st["node_tests"] / n, st["elem_tests"] / n, st["recheck_pairs"] / n))
