"""Latency of the host entry point for small batches (what a planner that calls IsFeasible one configuration at a time sees)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from klampt_b200 import synth
from klampt_b200.engine import Engine
w = synth.world_c2()
eng = Engine(w)
Q = synth.sample_configs(w.robot, 1 << 17, 3)
for n in (1, 32, 1024, 10000, 100000):
    q = np.ascontiguousarray(Q[:n]); out = np.empty(n, dtype=np.uint8)
    for _ in range(20):
        eng.feasible_batch(q, out=out)
    reps = 200 if n <= 10000 else 30
    t = time.perf_counter()
    for _ in range(reps):
        eng.feasible_batch(q, out=out)
    dt = (time.perf_counter() - t) / reps
    print("N=%6d: %.1f us per call -> %.3e cfg/s" % (n, dt * 1e6, n / dt))
A, B = synth.sample_edges(w.robot, lambda X: eng.feasible_batch(X), 64, 4)
for n in (1, 64):
    for _ in range(5):
        eng.edges_visible_batch(A[:n], B[:n], eps=0.01)
    t = time.perf_counter()
    for _ in range(30):
        eng.edges_visible_batch(A[:n], B[:n], eps=0.01)
    dt = (time.perf_counter() - t) / 30
    print("edges N=%3d: %.1f us per call -> %.3e edges/s" % (n, dt * 1e6, n / dt))
