"""Quick GPU look at one workload: parity vs oracle on a sample + device-resident throughput."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from oracle.oracle import OracleWorld, max_threads

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
w = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3}[which]()
t = time.time(); eng = Engine(w); print("engine build %.2fs" % (time.time() - t), eng.layout())
t = time.time(); orc = OracleWorld(w); print("oracle build %.2fs" % (time.time() - t))
Q = synth.sample_configs(w.robot, N, 2)
ns = min(N, 20000)
t = time.time(); want = orc.feasible_batch(Q[:ns], nthreads=0); dt = time.time() - t
print("oracle %d threads: %.0f cfg/s, feasible %.3f" % (max_threads(), ns / dt, want.mean()))
eng.set_option("collect_stats", 1)
if len(sys.argv) > 4:
    eng.set_option("pipeline", int(sys.argv[4]))
got = eng.feasible_batch(Q[:ns])
bad = np.nonzero(got != want)[0]
st = eng.stats()
print("mismatches:", len(bad), "per config: node %.1f elem %.1f recheck %.2f" % (st["node_tests"] / ns, st["elem_tests"] / ns, st["recheck_pairs"] / ns))
for i in bad[:10]:
    print("  cfg", i, "gpu", got[i], "oracle", want[i], "clearance", orc.distance(Q[i], 1.0, True))
eng.set_option("collect_stats", 0)
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    eng.set_option("chunk", int(sys.argv[3]))
if len(sys.argv) > 4:
    eng.set_option("pipeline", int(sys.argv[4]))
if len(sys.argv) > 5:
    eng.set_option("leaf_budget", int(sys.argv[5]))
dQ = torch.from_numpy(Q).cuda()
dout = torch.empty(N, dtype=torch.uint8, device="cuda")
eng.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    eng.feasible_batch_device(dQ, N, dout)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.feasible_batch_device(dQ, N, dout)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("%s: N=%d device-resident %.3f ms -> %.3e cfg/s; feasible %.3f" % (which, N, ms, N / ms * 1e3, dout.float().mean().item()))
t = time.time(); out = eng.feasible_batch(Q); dt = time.time() - t
print("host path (pageable): %.3f ms -> %.3e cfg/s" % (dt * 1e3, N / dt))
