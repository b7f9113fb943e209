"""GPU linear-BVH rebuild of a replaceable point cloud: update time, and query speed against the host-built SAH hierarchy on the same points."""
import sys, os, time, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.worldspec import GeomSpec, WorldSpec
from klampt_b200.engine import Engine
from scripts.gpu_grid import timed

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
ws = synth.world_c5(n_points=npts)                      # static: host SAH build at kb_finalize
cloud = [g for g in ws.geoms if g.kind == "cloud"][0]
t = time.time(); es = Engine(ws); t_static = time.time() - t
t = time.time(); eg = Engine(ws, options={"cloud_builder": 1}); t_gpu = time.time() - t
print("static cloud, %d points: kb_finalize with the host SAH build %.2f s, with cloud_builder = 1 (GPU) %.2f s" % (npts, t_static, t_gpu))
wd = WorldSpec()
gd = wd.add_geom(GeomSpec.dynamic_cloud(len(cloud.points), radius=0.0, margin=cloud.margin))
wd.terrains.append(gd)                                  # C5's cloud is a terrain
wd.robot = synth.make_arm6(wd)
t = time.time(); ed = Engine(wd); t_dyn = time.time() - t
P = np.ascontiguousarray(cloud.points)
ed.update_pointcloud(gd, P); ed.synchronize()
ts = []
for _ in range(5):
    t = time.time(); ed.update_pointcloud(gd, P); ed.synchronize(); ts.append(time.time() - t)
print("points %d: engine build static (host SAH) %.2f s, dynamic (reserved) %.2f s; GPU rebuild incl. H2D of the points: %.1f ms (best of 5: %.1f ms)"
      % (len(P), t_static, t_dyn, 1e3 * np.median(ts), 1e3 * min(ts)))
N = 200000
Q = synth.sample_configs(ws.robot, N, 5)
rs, rd, rg = es.feasible_batch(Q), ed.feasible_batch(Q), eg.feasible_batch(Q)
print("feasible: static %.4f dynamic %.4f mismatches %d; static built on the GPU: mismatches %d" % (rs.mean(), rd.mean(), int((rs != rd).sum()), int((rs != rg).sum())))
ds, dg = es.distance_batch(Q[:2000], upper_bound=0.3), eg.distance_batch(Q[:2000], upper_bound=0.3)
print("distance: max |diff| between the two hierarchies %.3g" % float(np.max(np.abs(ds - dg))))
dQ = torch.from_numpy(Q).cuda(); dout = torch.empty(N, dtype=torch.uint8, device="cuda")
for name, eng in (("SAH (host)", es), ("LBVH (GPU)", ed), ("LBVH static", eg)):
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_option("collect_stats", 1); eng.reset_stats(); eng.feasible_batch_device(dQ, N, dout); torch.cuda.synchronize(); st = eng.stats(); eng.set_option("collect_stats", 0)
    ms = timed(eng, dQ, N, dout)
    print("  %-11s boolean query %.3f ms / 200k -> %.3e cfg/s; per cfg: iter %.1f node %.0f elem %.0f" % (name, ms, N / ms * 1e3, st["node_iterations"] / N, st["node_tests"] / N, st["elem_tests"] / N))
