"""Where do the node tests of the arm6 self-collision check go?  Node tests per configuration with one self pair enabled at a time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from klampt_b200 import synth
from klampt_b200.worldspec import WorldSpec
from klampt_b200.engine import Engine

N = 100000
def run(edits, label):
    w = WorldSpec(); w.robot = synth.make_arm6(w); w.robot.self_collision_edits = edits
    eng = Engine(w); eng.set_option("collect_stats", 1)
    Q = synth.sample_configs(w.robot, N, 2)
    out = eng.feasible_batch(Q); st = eng.stats()
    print("%-10s items %2d  node %.1f elem %.1f  feasible %.3f" % (label, eng.layout()["items_per_config"], st["node_tests"] / N, st["elem_tests"] / N, out.mean()))
    eng.close()
run([], "all")
L = 7
pairs = [(i, j) for i in range(L) for j in range(i + 2, L)]
for a in pairs:
    run([(i, j, (i, j) == a) for (i, j) in pairs], "%d-%d" % a)
