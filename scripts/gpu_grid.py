"""Clearance-grid broad phase: identical results with the grids on / off, parity vs the oracle, items dropped, timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from oracle.oracle import OracleWorld

def timed(eng, dQ, N, dout, reps=3):
    for _ in range(2):
        eng.feasible_batch_device(dQ, N, dout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.feasible_batch_device(dQ, N, dout)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = (sys.argv[1:] or ["c1", "c2", "c5"]) if __name__ == "__main__" else []
for name in which:
    N = {"c1": 200000, "c2": 1000000, "c3": 200000, "c5": 200000}[name]
    w = {"c1": synth.world_c1, "c2": synth.world_c2, "c3": synth.world_c3, "c5": lambda: synth.world_c5(n_points=1000000)}[name]()
    t = time.time(); eng = Engine(w); tb = time.time() - t
    print("== %s: engine build %.2fs %s" % (name, tb, eng.layout()))
    Q = synth.sample_configs(w.robot, N, 2)
    dQ = torch.from_numpy(Q).cuda(); dout = torch.empty(N, dtype=torch.uint8, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    res = {}
    for on in (0, 1):
        eng.set_option("clear_grid", on)
        eng.set_option("collect_stats", 1); eng.reset_stats()
        eng.feasible_batch_device(dQ, N, dout); torch.cuda.synchronize()
        st = eng.stats(); res[on] = dout.cpu().numpy().copy()
        eng.set_option("collect_stats", 0)
        ms = timed(eng, dQ, N, dout)
        print("  grid=%d: %.3f ms -> %.3e cfg/s; feasible %.4f; per cfg: node %.1f elem %.1f recheck %.2f dropped %.2f of %d items"
              % (on, ms, N / ms * 1e3, res[on].mean(), st["node_tests"] / N, st["elem_tests"] / N, st["recheck_pairs"] / N, st.get("items_dropped", 0) / N, eng.layout()["items_per_config"]))
    print("  grid on vs off mismatches:", int((res[0] != res[1]).sum()))
    ns = 20000
    orc = OracleWorld(w)
    want = orc.feasible_batch(Q[:ns], nthreads=0)
    print("  vs oracle (%d): mismatches %d" % (ns, int((res[1][:ns] != want).sum())))
    eng.close()
