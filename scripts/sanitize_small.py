"""small end-to-end pass for compute-sanitizer (memcheck / racecheck): feasibility, edges, distance, pair queries"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from klampt_b200 import synth
from klampt_b200.engine import Engine
w = synth.world_c2(2, n_obstacles=20)
eng = Engine(w)
Q = synth.sample_configs(w.robot, 3000, 1)
f, pairs = eng.feasible_batch(Q, return_pairs=True)
A, B = Q[f == 1][:64], Q[f == 1][64:128]
v, n = eng.edges_visible_batch(A, B, eps=0.02)
d = eng.distance_batch(Q[:300], upper_bound=0.3, include_self=True)
w3 = synth.world_c3()
e3 = Engine(w3)
f3 = e3.feasible_batch(synth.sample_configs(w3.robot, 2000, 3))
w5 = synth.world_c5(n_points=20000, n_obstacles=20)
e5 = Engine(w5)
f5 = e5.feasible_batch(synth.sample_configs(w5.robot, 1500, 5))
d5 = e5.distance_batch(synth.sample_configs(w5.robot, 200, 5), upper_bound=0.5)
eng.set_option("clear_grid", 1)
fg = eng.feasible_batch(Q)
assert (fg == f).all()
cp, cc = eng.colliding_pairs_batch(Q[:500], max_pairs=8)
wf = synth.world_floating(n_obstacles=10)
ef = Engine(wf)
Qf = synth.sample_configs(wf.robot, 1500, 9)
ff = ef.feasible_batch(Qf)
vf, nf = ef.edges_visible_batch(Qf[ff == 1][:32], Qf[ff == 1][32:64], eps=0.05)
wb = synth.world_boxes(n_boxes=8, n_blobs=2)
eb = Engine(wb)
Qb = synth.sample_configs(wb.robot, 1500, 11)
fb = eb.feasible_batch(Qb)
db = eb.distance_batch(Qb[:200], upper_bound=0.3, include_self=True)
cpb, ccb = eb.colliding_pairs_batch(Qb[:300], max_pairs=8)
import copy
from klampt_b200.worldspec import GeomSpec
wd = copy.deepcopy(synth.world_c1())
gdyn = wd.add_geom(GeomSpec.dynamic_cloud(50000, radius=0.003, margin=0.001))
wd.objects.append((gdyn, synth.make_T(None, (0.1, 0.0, 0.2))))
wd.robot = synth.make_arm6(wd)
ed = Engine(wd)
rng = np.random.default_rng(3)
for n in (30000, 7, 0, 49999):
    ed.update_pointcloud(gdyn, rng.uniform([-0.8, -0.8, 0.1], [0.8, 0.8, 1.2], size=(n, 3)))
    fd = ed.feasible_batch(synth.sample_configs(wd.robot, 1500, 12))
ff32 = eng.feasible_batch(Q.astype(np.float32))
print("ok", f.mean(), v.mean(), float(d.min()), f3.mean(), f5.mean(), float(d5.min()))
