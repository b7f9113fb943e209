"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the fp64 oracle):
CPU leg pins the oracle, GPU leg checks the CUDA path through the C ABI against the same files."""
import os

import numpy as np
import pytest

from klampt_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
WORLDS = {"c1": lambda: synth.world_c1(), "c2_60": lambda: synth.world_c2(2, n_obstacles=60), "c3": lambda: synth.world_c3(),
          "boxes": lambda: synth.world_boxes(), "floating": lambda: synth.world_floating()}
EDGE_EPS = {"floating": 0.02, "boxes": 0.02}
# in-band mismatches tolerated per golden world: 0 = exact equality (mesh-mesh, margin 0); solids are distance-threshold elements
GOLDEN_BAND_ALLOWANCE = {"boxes": 4}


def load(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    spec = WORLDS[name]()
    Q = synth.sample_configs(spec.robot, int(g["n_cfg"]), int(g["seed"]))
    feas = np.unpackbits(g["feasible_bits"])[:len(Q)]
    return g, spec, Q, feas


@pytest.mark.parametrize("name", sorted(WORLDS))
def test_oracle_reproduces_golden(name):
    from oracle.oracle import OracleWorld
    g, spec, Q, feas = load(name)
    o = OracleWorld(spec)
    n = 1200
    assert np.array_equal(o.feasible_batch(Q[:n]), feas[:n])
    np.testing.assert_allclose(o.fk_batch(Q[:16]), g["fk"], rtol=0, atol=1e-13)
    if "edge_A" in g.files:
        m = 60
        vis, nchk = o.edges_visible_batch(g["edge_A"][:m], g["edge_B"][:m], eps=EDGE_EPS.get(name, 0.01))
        assert np.array_equal(vis, np.unpackbits(g["edge_visible"])[:m]) and np.array_equal(nchk, g["edge_nchecks"][:m])
    if "dist_env" in g.files:
        d, _ = o.distance_batch(Q[:50], upper_bound=0.5)
        np.testing.assert_allclose(d, g["dist_env"][:50], rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(WORLDS))
def test_gpu_matches_golden(name, built):
    from klampt_b200.engine import Engine
    g, spec, Q, feas = load(name)
    eng = Engine(spec)
    got = eng.feasible_batch(Q)
    if not np.array_equal(got, feas):             # mismatches are only legal inside the two-sided 1e-6 m band (tests/parity.py)
        from oracle.oracle import OracleWorld
        from parity import assert_bool_parity
        assert_bool_parity(got, feas, Q, OracleWorld(spec), max_bad=GOLDEN_BAND_ALLOWANCE.get(name, 0))
    np.testing.assert_allclose(eng.fk_batch(Q[:16]), g["fk"], rtol=0, atol=1e-12)
    if "edge_A" in g.files:
        vis, nchk = eng.edges_visible_batch(g["edge_A"], g["edge_B"], eps=EDGE_EPS.get(name, 0.01))
        # Floating / BallAndSocket midpoints go through device libm (sincos, acos, atan2): equal to ~1e-15, not bit for bit
        slack = 1 if name == "floating" else 0
        assert (vis != np.unpackbits(g["edge_visible"])[:len(vis)]).sum() <= slack
        assert (nchk != g["edge_nchecks"]).sum() <= slack
    if "dist_env" in g.files:
        n = int(g["n_dist"])
        d, pr = eng.distance_batch(Q[:n], upper_bound=0.5, include_self=False, return_pairs=True)
        np.testing.assert_allclose(d, g["dist_env"], rtol=1e-5, atol=1e-9)
    if "dist_all" in g.files:
        n = int(g["n_dist"])
        np.testing.assert_allclose(eng.distance_batch(Q[:n], upper_bound=0.25, include_self=True), g["dist_all"], rtol=1e-5, atol=1e-9)
