"""Distance query cost on the point-cloud world (C5) and the mesh world (C2): node / element tests per configuration and time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from oracle.oracle import OracleWorld

for name in sys.argv[1:] or ["c5", "c2"]:
    w = {"c5": lambda: synth.world_c5(n_points=2000000), "c2": synth.world_c2}[name]()
    eng = Engine(w)
    N = 100000
    Q = synth.sample_configs(w.robot, N, 5)
    dQ = torch.from_numpy(Q).cuda(); dd = torch.empty(N, dtype=torch.float64, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for ub in (0.5, 0.1):
        eng.set_option("collect_stats", 1); eng.reset_stats()
        eng.distance_batch_device(dQ, N, ub, False, dd); torch.cuda.synchronize()
        st = eng.stats(); eng.set_option("collect_stats", 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            eng.distance_batch_device(dQ, N, ub, False, dd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print("%s ub=%.1f: %.2f ms / 100k -> %.3e cfg/s; per cfg: node %.0f elem %.0f" % (name, ub, ms, N / ms * 1e3, st["node_tests"] / N, st["elem_tests"] / N))
    if name == "c5":
        orc = OracleWorld(w)
        ns = 2000
        do, _ = orc.distance_batch(Q[:ns], upper_bound=0.5, include_self=False)
        d = dd.cpu().numpy()[:ns] if ub == 0.5 else eng.distance_batch(Q[:ns], upper_bound=0.5)
        d = eng.distance_batch(Q[:ns], upper_bound=0.5)
        print("  max rel err vs oracle:", float(np.max(np.abs(d - do) / np.maximum(1e-9, np.abs(do)))))
