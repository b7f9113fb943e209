import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
w = synth.world_c5()
n = 200_000
Q = synth.sample_configs(w.robot, n, 3)
dQ = torch.from_numpy(Q).cuda(); out = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for wide in (1, 0):
    eng = Engine(w, options={"wide": wide, "cloud_builder": 0})
    eng.feasible_batch_device(dQ, n, out); eng.synchronize()
    t = time.perf_counter(); eng.feasible_batch_device(dQ, n, out); eng.synchronize(); dt = time.perf_counter() - t
    res[wide] = out.cpu().numpy().copy()
    eng.set_option("collect_stats", 1); eng.reset_stats()
    eng.feasible_batch_device(dQ, n, out); eng.synchronize()
    st = eng.stats()
    print("c5 boolean wide=%d: %.2f ms per %d; node tests %.0f, iterations %.1f, element tests %.1f, rechecks %.2f per configuration; %s" % (
        wide, 1e3 * dt, n, st["node_tests"] / n, st["node_iterations"] / n, st["elem_tests"] / n, st["recheck_pairs"] / n, eng.layout()))
    eng.close()
print("equal:", np.array_equal(res[0], res[1]), res[0].mean())
