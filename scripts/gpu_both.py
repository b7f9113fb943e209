"""Experiment: descend both trees at once on narrow frontiers (option both_limit); iterations / node tests / time."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from scripts.gpu_grid import timed

for name in sys.argv[1:] or ["c2", "c3"]:
    N = 1000000
    w = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3}[name]()
    eng = Engine(w)
    Q = synth.sample_configs(w.robot, N, 2)
    dQ = torch.from_numpy(Q).cuda(); dout = torch.empty(N, dtype=torch.uint8, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    ref = None
    for grid in (0,):
        for bl in (0,):
            eng.set_option("clear_grid", grid); eng.set_option("both_limit", bl)
            eng.set_option("collect_stats", 1); eng.reset_stats()
            eng.feasible_batch_device(dQ, N, dout); torch.cuda.synchronize()
            st = eng.stats(); r = dout.cpu().numpy().copy()
            if ref is None: ref = r
            eng.set_option("collect_stats", 0)
            ms = timed(eng, dQ, N, dout)
            print("%s grid=%d both_limit=%2d: %.3f ms; per cfg: iter %.1f node %.1f (%.1f lanes/iter) elem %.1f; mismatches vs base %d"
                  % (name, grid, bl, ms, st["node_iterations"] / N, st["node_tests"] / N, st["node_tests"] / max(1, st["node_iterations"]), st["elem_tests"] / N, int((r != ref).sum())))
    eng.close()
