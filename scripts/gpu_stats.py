"""traversal statistics per configuration (node tests, warp iterations, element tests, fp64 rechecks) for a workload, wide vs binary"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3}[which]()
n = 200_000
Q = synth.sample_configs(w.robot, n, 3)
dQ = torch.from_numpy(Q).cuda(); out = torch.empty(n, dtype=torch.uint8, device="cuda")
for wide in (1, 0):
    eng = Engine(w, options={"wide": wide})
    eng.set_option("collect_stats", 1); eng.reset_stats()
    eng.feasible_batch_device(dQ, n, out); eng.synchronize()
    st = eng.stats()
    print("%s wide=%d: node tests %.0f, iterations %.1f (%.1f tests / iteration), element tests %.1f, rechecks %.2f per configuration; layout %s" % (
        which, wide, st["node_tests"] / n, st["node_iterations"] / n, st["node_tests"] / max(1, st["node_iterations"]), st["elem_tests"] / n, st["recheck_pairs"] / n, eng.layout()))
    eng.close()
