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


def test_c4_one_hundred_thousand_edges(built):
    """BASELINE config 4 at a size the oracle still finishes in seconds: visibility AND the sequential checker's nchecks, bit for bit"""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2()
    eng, orc = Engine(w), OracleWorld(w)
    A, B = synth.sample_edges(w.robot, lambda Q: eng.feasible_batch(Q), 100_000, 44)
    vis, n = eng.edges_visible_batch(A, B, eps=0.01)
    ovis, on = orc.edges_visible_batch(A, B, eps=0.01, nthreads=0)
    assert np.array_equal(vis, ovis) and np.array_equal(n, on)
    bits = eng.edges_visible_batch_bits(A, B, eps=0.01)
    assert np.array_equal(np.unpackbits(bits, bitorder="little")[:len(A)], vis)
    assert 0.4 < vis.mean() < 0.7


def test_c5_five_million_point_cloud(built):
    """BASELINE config 5 at full cloud size (5M points): collide bit and min distance (upper bound 0.5 m) of 100k configurations
    against the oracle -- booleans under the two-sided band rule, distances within 1e-5 relative"""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    from parity import assert_bool_parity
    w = synth.world_c5()
    eng, orc = Engine(w, options={"cloud_builder": 1}), OracleWorld(w)
    n = 100_000
    Q = synth.sample_configs(w.robot, n, 55)
    got = eng.feasible_batch(Q)
    want = orc.feasible_batch(Q, nthreads=0)
    assert_bool_parity(got, want, Q, orc, max_bad=n // 1000)       # point spheres with a margin: a distance threshold in fp32 + fp64 recheck
    assert 0.2 < got.mean() < 0.9
    d = eng.distance_batch(Q, upper_bound=0.5, include_self=False)
    od, _ = orc.distance_batch(Q, upper_bound=0.5, include_self=False, nthreads=0)
    np.testing.assert_allclose(d, od, rtol=1e-5, atol=1e-9)
    # consistency of the two query kinds: environment clearance <= 0 exactly where the environment collides (self pairs aside)
    env_hit = d <= 0
    assert not (env_hit & (got == 1)).any()
    bits = eng.feasible_batch_bits(Q)
    assert np.array_equal(np.unpackbits(bits, bitorder="little")[:n], got)


def test_c6_depth_images_at_full_world_size(built):
    """the ray-cast row at BASELINE's world sizes: a 640 x 480 image of the full C2 world (200 obstacles, 505 k triangles) and one of
    the 5 M-point cloud world, every pixel against the oracle's per-body loop: same body, distance to 1e-9"""
    import math
    from klampt_b200 import sensing
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    for world, opts in ((synth.world_c2(), None), (synth.world_c5(), {"cloud_builder": 1})):
        eng, orc = Engine(world, options=opts), OracleWorld(world)
        q = synth.sample_configs(world.robot, 1, 77)[0]
        eye = np.array([3.2 * math.cos(0.4), 3.2 * math.sin(0.4), 1.3])
        fwd = np.array([0.0, 0.0, 0.5]) - eye
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, [0.0, 0.0, 1.0])
        right /= np.linalg.norm(right)
        cam = sensing.CameraSensor(640, 480, zmin=0.1, zmax=8.0, Tsensor=synth.make_T(np.stack([right, np.cross(fwd, right), fwd], axis=1), eye))
        rays, _, _ = cam.rays()
        ids, dist, elem = eng.raycast_batch(q, rays)
        oid, od, oel = orc.raycast_batch(q, rays, nthreads=0)
        assert np.array_equal(ids, oid)
        hit = ids >= 0
        assert 0.3 < hit.mean() <= 1.0
        np.testing.assert_allclose(dist[hit], od[hit], rtol=1e-9, atol=1e-12)
        assert (elem == oel)[hit].mean() > 0.999                      # ties across a shared edge / coincident points aside
        depth, idimg = cam.simulate(eng, q)                            # the device-built rays give the same picture
        assert (idimg.reshape(-1) == ids).mean() > 0.9999             # (host-normalised directions differ from the device's in the last bit)
