"""Regenerates tests/golden/*.npz.

The reference cannot be built or imported here (its arithmetic for this path lives in KrisLibrary, absent from the
reference tree), and its own tests hold no vectors for FK / collision / distance / edge visibility, so
these fixtures are produced by the CPU oracle (oracle/kb_oracle.c, fp64) on the seeded synthetic worlds, next to
closed-form answers where they exist.  They pin the oracle against regressions (tests/test_golden.py, CPU) and give
the GPU parity tests a committed target that does not depend on rebuilding the oracle on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from klampt_b200 import synth          # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EDGE_EPS = {"floating": 0.02, "boxes": 0.02}


def world_fixture(name, spec, n_cfg, n_edges, n_dist, seed):
    o = OracleWorld(spec)
    Q = synth.sample_configs(spec.robot, n_cfg, seed)
    feas, pairs = o.feasible_batch(Q, want_pairs=True)
    out = dict(seed=np.int64(seed), n_cfg=np.int64(n_cfg), feasible_bits=np.packbits(feas), fk=o.fk_batch(Q[:16]),
               limits_ok=np.packbits(np.array([o.check_joint_limits(q) for q in Q], dtype=np.uint8)))
    if n_edges:
        A, B = synth.sample_edges(spec.robot, lambda X: o.feasible_batch(X), n_edges, seed)
        vis, nchk = o.edges_visible_batch(A, B, eps=EDGE_EPS.get(name, 0.01))
        out.update(edge_seed=np.int64(seed), n_edges=np.int64(n_edges), edge_A=A, edge_B=B, edge_visible=np.packbits(vis), edge_nchecks=nchk.astype(np.int32))
    if n_dist:
        d, dp = o.distance_batch(Q[:n_dist], upper_bound=0.5, include_self=False)
        ds, _ = o.distance_batch(Q[:n_dist], upper_bound=0.25, include_self=True)
        out.update(n_dist=np.int64(n_dist), dist_env=d, dist_env_pair=dp.astype(np.int32), dist_all=ds)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "feasible %.3f" % feas.mean(), {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    world_fixture("c1", synth.world_c1(), 4000, 300, 300, 1)
    world_fixture("c2_60", synth.world_c2(2, n_obstacles=60), 4000, 0, 200, 2)
    world_fixture("c3", synth.world_c3(), 4000, 150, 0, 3)
    world_fixture("boxes", synth.world_boxes(), 4000, 100, 200, 8)          # solid box primitives (links, objects, terrain)
    world_fixture("floating", synth.world_floating(), 4000, 150, 100, 7)    # Floating / BallAndSocket joints in the edge metric
