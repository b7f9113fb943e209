"""times feasibility and distance separately on one workload (device-resident)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
which = sys.argv[1]; N = int(sys.argv[2])
w = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3, "c5": synth.world_c5}[which]()
eng = Engine(w)
Q = synth.sample_configs(w.robot, N, 2)
dQ = torch.from_numpy(Q).cuda(); out = torch.empty(N, dtype=torch.uint8, device="cuda"); dd = torch.empty(N, dtype=torch.float64, device="cuda")
s = torch.cuda.Stream(); eng.set_stream(s.cuda_stream)
def timeit(fn, reps=3):
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): fn()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
eng.set_option("collect_stats", 1); eng.reset_stats()
eng.feasible_batch_device(dQ, N, out); st = eng.stats()
print("feasible: node %.1f elem %.1f recheck %.2f per cfg" % (st["node_tests"]/N, st["elem_tests"]/N, st["recheck_pairs"]/N))
eng.reset_stats(); eng.distance_batch_device(dQ, N, 0.5, False, dd); st = eng.stats()
print("distance(ub=0.5): node %.1f elem %.1f per cfg" % (st["node_tests"]/N, st["elem_tests"]/N))
eng.set_option("collect_stats", 0)
t = timeit(lambda: eng.feasible_batch_device(dQ, N, out)); print("%s feasible: %.3f ms -> %.3e cfg/s" % (which, t, N/t*1e3))
for ub in (0.5, 0.1, float("inf")):
    t = timeit(lambda: eng.distance_batch_device(dQ, N, ub, False, dd)); print("%s distance ub=%s: %.3f ms -> %.3e cfg/s" % (which, ub, t, N/t*1e3))
if which != "c5":
    t = timeit(lambda: eng.distance_batch_device(dQ, N, 0.5, True, dd)); print("%s distance incl. self ub=0.5: %.3f ms -> %.3e cfg/s" % (which, t, N/t*1e3))
