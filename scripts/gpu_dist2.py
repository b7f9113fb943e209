"""C5 distance kernel: time, traversal statistics and parity for a given cloud leaf size (argv[1], default 8)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from oracle.oracle import OracleWorld

leaf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
w = synth.world_c5()
t0 = time.time()
eng = Engine(w, options={"cloud_leaf": leaf})
print("leaf %d: finalize %.1f s, layout %s" % (leaf, time.time() - t0, eng.layout()))
Q = synth.sample_configs(w.robot, n, 55)
dQ = torch.from_numpy(Q).cuda()
dd = torch.empty(n, dtype=torch.float64, device="cuda")
eng.distance_batch_device(dQ, n, 0.5, False, dd); eng.synchronize()
ts = []
for _ in range(3):
    t = time.perf_counter(); eng.distance_batch_device(dQ, n, 0.5, False, dd); eng.synchronize(); ts.append(time.perf_counter() - t)
print("leaf %d: distance %d configs: %.2f ms  -> %.3g cfg/s" % (leaf, n, 1e3 * min(ts), n / min(ts)))
eng.set_option("collect_stats", 1); eng.reset_stats()
eng.distance_batch_device(dQ, n, 0.5, False, dd); eng.synchronize()
st = eng.stats()
print("   per config: node tests %.0f, element tests %.0f, fp64 evaluations %.1f, warp iterations %.0f" %
      (st["node_tests"] / n, st["elem_tests"] / n, st["recheck_pairs"] / n, st["node_iterations"] / n))
eng.set_option("collect_stats", 0)
m = 3000
orc = OracleWorld(w, variant="fast")
od, _ = orc.distance_batch(Q[:m], upper_bound=0.5)
d = dd[:m].cpu().numpy()
err = np.abs(d - od) / np.maximum(1e-9, np.abs(od))
print("   parity on %d: max rel err %.3g, max abs err %.3g" % (m, err.max(), np.abs(d - od).max()))
