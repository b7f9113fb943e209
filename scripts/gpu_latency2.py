"""Where a small kb_feasible_batch call spends its time: whole call (host clock), traversal kernel (CUDA events), the rest."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from klampt_b200 import synth
from klampt_b200.engine import Engine
from klampt_b200._capi import check
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = synth.world_c2() if which == "c2" else synth.world_c1()
eng = Engine(w)
Q = synth.sample_configs(w.robot, 1 << 17, 3)
for n in (1, 1024, 10000, 20000, 32768, 50000, 100000):
    q = np.ascontiguousarray(Q[:n]); out = np.empty(n, dtype=np.uint8)
    qp, op = C.c_void_p(q.ctypes.data), C.c_void_p(out.ctypes.data)
    for _ in range(30):
        check(eng.lib.kb_feasible_batch(eng.h, qp, n, op, None))
    reps = 300
    t = time.perf_counter()
    for _ in range(reps):
        check(eng.lib.kb_feasible_batch(eng.h, qp, n, op, None))
    dt = (time.perf_counter() - t) / reps
    eng.set_option("time_kernels", 1); eng.reset_stats()
    for _ in range(50):
        check(eng.lib.kb_feasible_batch(eng.h, qp, n, op, None))
    st = eng.stats(); eng.set_option("time_kernels", 0)
    print("%s N=%6d: %.1f us per call (%.3e cfg/s); traversal kernel %.1f us, gpu_ms per call %.1f us" % (which, n, dt * 1e6, n / dt, 1e3 * st["traverse_ms"] / max(1, st["traverse_launches"]), 1e3 * st["gpu_ms"] / 50))
