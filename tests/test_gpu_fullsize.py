"""Parity at BASELINE.json's bench sizes: the CPU oracle is fast enough on the GPU box's host cores (about 1e6
configurations/s all-core) to check the whole 1M-configuration C2 / C3 batches, not just a sample, plus size-independent
properties (idempotence, permutation invariance, consistency of the three query kinds)."""
import numpy as np
import pytest

from klampt_b200 import synth

pytestmark = pytest.mark.gpu


def _check(world, n, seed, built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    eng, orc = Engine(world), OracleWorld(world)
    Q = synth.sample_configs(world.robot, n, seed)
    got = eng.feasible_batch(Q)
    want = orc.feasible_batch(Q, nthreads=0)
    from parity import assert_bool_parity
    assert_bool_parity(got, want, Q, orc, max_bad=0)   # mesh-mesh, margin 0: exact equality with the fp64 oracle
    return eng, orc, Q, got


def test_c2_one_million_configurations(built):
    w = synth.world_c2()
    eng, orc, Q, got = _check(w, 1_000_000, 1234, built)
    assert 0.40 < got.mean() < 0.48
    # idempotent and independent of batch order / batch boundaries
    perm = np.random.default_rng(1).permutation(200_000)
    assert np.array_equal(eng.feasible_batch(Q[:200_000][perm]), got[:200_000][perm])
    assert np.array_equal(np.concatenate([eng.feasible_batch(Q[:77_777]), eng.feasible_batch(Q[77_777:200_000])]), got[:200_000])
    # the three query kinds agree: clearance > 0 everywhere <=> feasible (given limits hold, which they do for uniform samples)
    d = eng.distance_batch(Q[:20_000], upper_bound=0.05, include_self=True)
    assert np.array_equal(d > 0, got[:20_000] == 1)
    # an edge is visible only if its midpoint is feasible
    A, B = Q[:20_000:2], Q[1:20_000:2]
    vis = eng.edges_visible_batch(A, B, eps=0.05, return_nchecks=False)
    mid = eng.feasible_batch(0.5 * A + 0.5 * B)
    assert not (vis.astype(bool) & (mid == 0)).any()


def test_c3_one_million_configurations(built):
    eng, orc, Q, got = _check(synth.world_c3(), 1_000_000, 4321, built)
    assert 0.30 < got.mean() < 0.42


def test_c4_edges_against_oracle(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2()
    eng, orc = Engine(w), OracleWorld(w)
    A, B = synth.sample_edges(w.robot, lambda Q: eng.feasible_batch(Q), 20_000, 4)
    vis, n = eng.edges_visible_batch(A, B, eps=0.01)
    ovis, on = orc.edges_visible_batch(A, B, eps=0.01, nthreads=0)
    assert np.array_equal(vis, ovis) and np.array_equal(n, on)
    assert 0.4 < vis.mean() < 0.7 and n.max() >= 255
