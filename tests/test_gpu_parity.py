"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): feasible / visible booleans bit-exact outside a 1e-6 m clearance band,
distances within 1e-5 relative.  A boolean mismatch is tolerated only if the oracle's own clearance for that
configuration is inside the band."""
import numpy as np
import pytest

from klampt_b200 import synth
from klampt_b200.worldspec import GeomSpec, WorldSpec

pytestmark = pytest.mark.gpu

from parity import BAND, assert_bool_parity, assert_geom_bool_parity


@pytest.fixture(scope="module")
def c1(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c1()
    return w, Engine(w), OracleWorld(w)


@pytest.fixture(scope="module")
def c2small(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2(2, n_obstacles=60)
    return w, Engine(w), OracleWorld(w)


@pytest.fixture(scope="module")
def c3(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c3()
    return w, Engine(w), OracleWorld(w)


def test_fk_matches_oracle(c1):
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 512, 11)
    T = eng.fk_batch(Q)
    To = orc.fk_batch(Q)
    assert T.shape == To.shape
    np.testing.assert_allclose(T, To, rtol=0, atol=1e-12)


def test_fk_dualarm_branching_tree(c3):
    w, eng, orc = c3
    Q = synth.sample_configs(w.robot, 256, 12)
    np.testing.assert_allclose(eng.fk_batch(Q), orc.fk_batch(Q), rtol=0, atol=1e-12)


def test_feasible_c1(c1):
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 10000, 1)
    got, pairs = eng.feasible_batch(Q, return_pairs=True)
    want, opairs = orc.feasible_batch(Q, want_pairs=True)
    assert_bool_parity(got, want, Q, orc)
    assert 0.2 < got.mean() < 0.9
    # the reported pair must be a pair the mask enables, and -1,-1 exactly for feasible / limit-violating configurations
    mask = eng.pair_mask()
    hit = pairs[:, 0] >= 0
    assert not (hit & (got == 1)).any()
    for a, b in pairs[hit][:500]:
        assert mask[a, b] or mask[b, a]


def test_joint_limits_closed_interval(c1):
    w, eng, orc = c1
    r = w.robot
    Q = np.tile(0.5 * (r.qmin + r.qmax), (6, 1))
    Q[0, 2] = r.qmax[2]                      # on the bound: allowed
    Q[1, 2] = np.nextafter(r.qmax[2], 10)    # one ulp outside: infeasible
    Q[2, 3] = r.qmin[3]
    Q[3, 3] = np.nextafter(r.qmin[3], -10)
    Q[4, 0] = 1e-9                           # welded base joint has qmin=qmax=0
    got = eng.feasible_batch(Q)
    want = orc.feasible_batch(Q)
    assert list(got) == list(want)
    assert got[1] == 0 and got[3] == 0 and got[4] == 0


def test_feasible_c2_small(c2small):
    w, eng, orc = c2small
    Q = synth.sample_configs(w.robot, 20000, 2)
    got = eng.feasible_batch(Q)
    want = orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc)


def test_feasible_c3_self_collision(c3):
    w, eng, orc = c3
    Q = synth.sample_configs(w.robot, 20000, 3)
    got = eng.feasible_batch(Q)
    want = orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc)
    assert 0.05 < got.mean() < 0.95


def test_empty_and_ragged_batches(c1):
    w, eng, orc = c1
    assert eng.feasible_batch(np.zeros((0, w.robot.L))).shape == (0,)
    for n in (1, 7, 33, 4097):
        Q = synth.sample_configs(w.robot, n, 40 + n)
        assert (eng.feasible_batch(Q) == orc.feasible_batch(Q)).all()


def test_small_batches_plain_captured_and_replayed(c1):
    """the small-batch path of kb_feasible_batch: a size is run plainly the first time, captured into a CUDA graph the second time and
    replayed afterwards -- every call with different rows, all equal to the oracle; many distinct sizes (more than the graph cache
    holds) stay correct; the zero-copy sizes (<= 64) and the copied ones (> 64) both"""
    w, eng, orc = c1
    for n in (1, 5, 64, 65, 300, 5000):
        for rep in range(4):
            Q = synth.sample_configs(w.robot, n, 1000 * n + rep)
            assert np.array_equal(eng.feasible_batch(Q), orc.feasible_batch(Q)), (n, rep)
    for n in range(2, 42):                                  # 40 distinct sizes, each seen twice: the cache (32 entries) is recycled
        for rep in range(2):
            Q = synth.sample_configs(w.robot, n, 7 * n + rep)
            assert np.array_equal(eng.feasible_batch(Q), orc.feasible_batch(Q)), (n, rep)


def test_idempotent_and_order_independent(c1):
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 5000, 5)
    a = eng.feasible_batch(Q)
    b = eng.feasible_batch(Q)
    assert (a == b).all()
    perm = np.random.default_rng(0).permutation(len(Q))
    c = eng.feasible_batch(Q[perm])
    assert (c == a[perm]).all()


def test_edges_c1(c1):
    w, eng, orc = c1
    A, B = synth.sample_edges(w.robot, lambda Q: orc.feasible_batch(Q), 1500, 4)
    vis, nchk = eng.edges_visible_batch(A, B, eps=0.01)
    ovis, onchk = orc.edges_visible_batch(A, B, eps=0.01)
    assert (vis == ovis).all()
    assert (nchk == onchk).all()
    assert 0.05 < vis.mean() < 0.95


def test_edges_weighted_metric(c1):
    w, eng, orc = c1
    A, B = synth.sample_edges(w.robot, lambda Q: orc.feasible_batch(Q), 300, 6)
    wts = np.array([1.0, 2.0, 1.5, 1.0, 0.5, 0.25, 0.1])
    vis, nchk = eng.edges_visible_batch(A, B, eps=0.02, weights=wts)
    ovis, onchk = orc.edges_visible_batch(A, B, eps=0.02, weights=wts)
    assert (vis == ovis).all() and (nchk == onchk).all()


def test_distance_c1(c1):
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 1500, 7)
    d, pairs = eng.distance_batch(Q, upper_bound=0.5, include_self=False, return_pairs=True)
    do, _ = orc.distance_batch(Q, upper_bound=0.5, include_self=False)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
    assert (d <= 0.5).all() and (d >= 0).all()
    assert ((pairs[:, 0] < 0) == (d >= 0.5)).all()


def test_distance_with_self_pairs(c3):
    w, eng, orc = c3
    Q = synth.sample_configs(w.robot, 600, 8)
    d = eng.distance_batch(Q, upper_bound=0.3, include_self=True)
    do, _ = orc.distance_batch(Q, upper_bound=0.3, include_self=True)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)


def test_margins_add_to_threshold(built):
    """A,B collide iff dist(A,B) <= margin_A + margin_B (reference Cpp/docs/Manual-Geometry.md:17)."""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c1()
    for gi, _ in w.objects:
        w.geoms[gi].margin = 0.01
    for gi in w.robot.link_geom:
        w.geoms[gi].margin = 0.005
    eng, orc = Engine(w), OracleWorld(w)
    Q = synth.sample_configs(w.robot, 6000, 9)
    got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc, max_bad=len(Q) // 1000)     # fp32 distances with margins: in-band mismatches are legal
    w0 = synth.world_c1()
    base = OracleWorld(w0).feasible_batch(Q)
    assert want.sum() < base.sum()            # margins make the world strictly tighter


def test_geometry_pair_queries(built):
    """Geometry3D.collides / withinDistance / distance on two unit cubes (tests/objects/cube.off solid)."""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    v, t = synth.unit_cube()
    w = WorldSpec()
    ga = w.add_geom(GeomSpec.mesh(v, t))
    gb = w.add_geom(GeomSpec.mesh(v, t))
    gs = w.add_geom(GeomSpec.sphere([0.5, 0.5, 0.5], 0.25))
    w.robot = synth.make_planar_nR(w, 2)
    eng, orc = Engine(w), OracleWorld(w)
    rng = np.random.default_rng(3)
    N = 400
    Ta = np.tile(synth.IDENTITY12, (N, 1))
    Tb = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-1.6, 1.6, size=3)) for _ in range(N)])
    got = eng.geom_collides_batch(ga, Ta, gb, Tb)
    want = np.array([orc.geom_collides(ga, Ta[i], gb, Tb[i]) for i in range(N)])
    assert (got == want).all()
    d = eng.geom_distance_batch(ga, Ta, gb, Tb)
    do = np.array([orc.geom_distance(ga, Ta[i], gb, Tb[i]) for i in range(N)])
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-12)
    assert ((d == 0) == want).all()
    wd = eng.geom_collides_batch(ga, Ta, gb, Tb, tol=0.2)
    assert (wd == (do <= 0.2)).all()
    # sphere primitive vs mesh: distance = dist(centre, surface) - r, negative when the surface cuts the ball
    ds = eng.geom_distance_batch(gs, Ta, gb, Tb)
    dso = np.array([orc.geom_distance(gs, Ta[i], gb, Tb[i]) for i in range(N)])
    np.testing.assert_allclose(ds, dso, rtol=1e-5, atol=1e-12)
    # analytic: cubes offset by 1.5 along x -> gap 0.5; offset 0.5 -> faces cross
    Tx = np.tile(synth.IDENTITY12, (2, 1)); Tx[0, 9] = 1.5; Tx[1, 9] = 0.5; Tx[1, 10] = 0.25; Tx[1, 11] = 0.25
    dd = eng.geom_distance_batch(ga, Ta[:2], gb, Tx)
    assert abs(dd[0] - 0.5) < 1e-12 and dd[1] == 0.0


@pytest.fixture(scope="module")
def c5small(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c5(n_points=60000, n_obstacles=60)
    return w, Engine(w), OracleWorld(w)


def test_point_cloud_collide_with_margin(c5small):
    """robot meshes vs point cloud (points as spheres of radius = margin 5 mm): CollisionPointCloud semantics"""
    w, eng, orc = c5small
    Q = synth.sample_configs(w.robot, 8000, 51)
    got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc, max_bad=len(Q) // 1000)     # fp32 distances with margins: in-band mismatches are legal
    assert 0.2 < got.mean() < 0.9


def test_point_cloud_distance(c5small):
    """mesh - cloud distance is unsigned geometric distance minus margins, capped at upperBound"""
    w, eng, orc = c5small
    Q = synth.sample_configs(w.robot, 1200, 52)
    d, pairs = eng.distance_batch(Q, upper_bound=0.5, include_self=False, return_pairs=True)
    do, po = orc.distance_batch(Q, upper_bound=0.5, include_self=False)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
    assert (d >= -0.005 - 1e-12).all()            # margin 5 mm is subtracted: touching points report -0.005
    assert ((pairs[:, 1] == 0) | (pairs[:, 1] == -1)).all()    # the cloud is terrain 0


def test_point_cloud_with_radii_and_cloud_cloud(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(8)
    w = WorldSpec()
    ga = w.add_geom(GeomSpec.cloud(rng.normal(size=(3000, 3)) * 0.3, rng.uniform(0.0, 0.03, size=3000)))
    gb = w.add_geom(GeomSpec.cloud(rng.normal(size=(500, 3)) * 0.2, None, margin=0.01))
    gm = w.add_geom(GeomSpec.mesh(*synth.blob_mesh(rng, 2, 0.25)))
    w.robot = synth.make_planar_nR(w, 2)
    eng, orc = Engine(w), OracleWorld(w)
    N = 300
    Ta = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-0.2, 0.2, size=3)) for _ in range(N)])
    Tb = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-1.2, 1.2, size=3)) for _ in range(N)])
    for g1, g2 in ((ga, gb), (ga, gm), (gm, gb)):
        d = eng.geom_distance_batch(g1, Ta, g2, Tb)
        do = np.array([orc.geom_distance(g1, Ta[i], g2, Tb[i]) for i in range(N)])
        np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-12)
        c = eng.geom_collides_batch(g1, Ta, g2, Tb)
        co = np.array([orc.geom_collides(g1, Ta[i], g2, Tb[i]) for i in range(N)])
        assert (c == co).all()
        assert (c == (do <= 0)).all()


def test_multi_chunk_batches(c1):
    """a batch larger than the per-launch chunk is split across launches (and fp64 rechecks parked in one launch are
    resolved before it ends)"""
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 5000, 71)
    want = orc.feasible_batch(Q)
    eng.set_option("chunk", 512)
    try:
        got, pairs = eng.feasible_batch(Q, return_pairs=True)
        assert (got == want).all()
        A, B = Q[want == 1][:300], Q[want == 1][300:600]
        vis, n = eng.edges_visible_batch(A, B, eps=0.02)
        ovis, on = orc.edges_visible_batch(A, B, eps=0.02)
        assert (vis == ovis).all() and (n == on).all()
        d = eng.distance_batch(Q[:1500], upper_bound=0.4)
        do, _ = orc.distance_batch(Q[:1500], upper_bound=0.4)
        np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
    finally:
        eng.set_option("chunk", 1 << 20)
    with pytest.raises(Exception):
        eng.set_option("chunk", 3)
    with pytest.raises(Exception):
        eng.set_option("no_such_option", 1)


def test_split_pipeline_gives_identical_results(c2small, c3):
    """option pipeline=1 (node kernel -> leaf-pair list -> leaf kernel -> requeue) against the oracle and the fused kernel,
    including a tiny leaf budget that forces most colliding configurations through the requeue path"""
    for (w, eng, orc), seed in ((c2small, 81), (c3, 82)):
        Q = synth.sample_configs(w.robot, 12000, seed)
        fused, fpairs = eng.feasible_batch(Q, return_pairs=True)
        want = orc.feasible_batch(Q)
        try:
            for budget in (64, 2):
                eng.set_option("pipeline", 1)
                eng.set_option("leaf_budget", budget)
                got, pairs = eng.feasible_batch(Q, return_pairs=True)
                assert (got == fused).all()
                assert_bool_parity(got, want, Q, orc)
                assert ((pairs[:, 0] >= 0) == (fpairs[:, 0] >= 0)).all()
        finally:
            eng.set_option("pipeline", 0)
            eng.set_option("leaf_budget", 64)


def test_prismatic_spin_driver_and_custom_mask(built):
    """the less common corners of the data model in one world: prismatic links, a branching tree, a Spin joint (metric and
    interpolation take the short arc), an affine driver limit, a self-collision edit and a user pair mask"""
    from klampt_b200.engine import Engine
    from klampt_b200.worldspec import DriverSpec, JOINT_SPIN, JOINT_NORMAL, JOINT_WELD, PRISMATIC
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(17)
    w = WorldSpec()
    v, t = synth.box_mesh([-1.5, -1.5, -0.4], [1.5, 1.5, -0.3], div=3)
    w.terrains.append(w.add_geom(GeomSpec.mesh(v, t)))
    for _ in range(8):
        d = rng.uniform(0.05, 0.2, size=3)
        bv, bt = synth.box_mesh(-d, d, div=2)
        w.objects.append((w.add_geom(GeomSpec.mesh(bv, bt)), synth.make_T(synth._random_rotation(rng), rng.uniform(-1.2, 1.2, size=3))))
    r = synth.make_planar_nR(w, 6, 0.35)
    r.parents = np.array([-1, 0, 1, 1, 3, 4], dtype=np.int32)           # branch at link 1
    r.linktype[2] = PRISMATIC
    r.axis[2] = [1.0, 0.0, 0.0]
    r.qmin[:] = [-np.inf, -2.0, -0.3, -2.0, -2.0, -2.0]
    r.qmax[:] = [np.inf, 2.0, 0.4, 2.0, 2.0, 2.0]
    r.joint_type = np.array([JOINT_SPIN, JOINT_NORMAL, JOINT_NORMAL, JOINT_NORMAL, JOINT_NORMAL, JOINT_NORMAL], dtype=np.uint8)
    r.drivers.append(DriverSpec(links=[3, 4], scale=[1.0, 2.0], offset=[0.1, 0.0], qmin=-0.8, qmax=0.8))
    r.self_collision_edits += [(0, 2, False), (2, 3, False), (2, 5, True)]      # siblings 2 and 3 share their joint origin
    w.robot = r
    n = w.num_ids()
    base = OracleWorld(w).pair_mask()
    mask = base.copy()
    lid = w.robot_link_id(5)
    mask[lid, w.rigid_object_id(0)] = 0
    mask[w.rigid_object_id(0), lid] = 0                                   # IgnoreCollisions(link 5, object 0)
    mask[w.robot_link_id(4), w.rigid_object_id(1)] = 0                     # one direction only: still enabled via the OR
    w.pair_mask = mask
    eng, orc = Engine(w), OracleWorld(w)
    assert np.array_equal(eng.pair_mask(), orc.pair_mask()) and eng.num_ids() == n
    Q = rng.uniform([-7, -2.1, -0.32, -2.1, -2.1, -2.1], [7, 2.1, 0.42, 2.1, 2.1, 2.1], size=(20000, 6))
    np.testing.assert_allclose(eng.fk_batch(Q[:200]), orc.fk_batch(Q[:200]), rtol=0, atol=1e-12)
    got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc, max_bad=max(2, len(Q) // 1000))   # distance-threshold elements: in-band mismatches are legal
    lim = np.array([orc.check_joint_limits(q) for q in Q[:2000]])
    assert 0.2 < lim.mean() < 0.9 and not got[:2000][~lim].any()
    ok = Q[want == 1]
    assert len(ok) >= 400
    A, B = ok[:200], ok[200:400]
    vis, nchk = eng.edges_visible_batch(A, B, eps=0.02)
    ovis, on = orc.edges_visible_batch(A, B, eps=0.02)
    assert (vis == ovis).all() and (nchk == on).all()
    d = eng.distance_batch(Q[:500], upper_bound=0.3, include_self=True)
    do, _ = orc.distance_batch(Q[:500], upper_bound=0.3, include_self=True)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)


def test_all_colliding_pairs(c1):
    """kb_colliding_pairs_batch lists every colliding id pair (the per-pair constraints of SingleRobotCSpace::Init evaluated
    together); checked against explicit per-pair oracle queries with set semantics"""
    from oracle.oracle import colliding_pairs
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 400, 91)
    Q[5, 2] = w.robot.qmax[2] + 1.0                                    # limits fail -> count -1
    pairs, count = eng.colliding_pairs_batch(Q, max_pairs=8)
    feas = eng.feasible_batch(Q)
    multi = 0
    for i in range(len(Q)):
        want = colliding_pairs(orc, Q[i])
        if want is None:
            assert count[i] == -1 and (pairs[i] == -1).all()
            continue
        got = {(int(a), int(b)) for a, b in pairs[i] if a >= 0}
        assert count[i] == len(want) and got == want, (i, got, want)
        assert (count[i] == 0) == bool(feas[i])
        multi += len(want) > 1
    assert multi > 5                                                  # the early-exit query would have reported one of these only
    # max_pairs smaller than the number found: count still reports all, the stored prefix is a subset
    p1, c1_ = eng.colliding_pairs_batch(Q, max_pairs=1)
    assert np.array_equal(c1_, count)
    for i in np.nonzero(count > 0)[0][:50]:
        assert (int(p1[i, 0, 0]), int(p1[i, 0, 1])) in {(int(a), int(b)) for a, b in pairs[i] if a >= 0}


def test_clearance_grid_option_changes_nothing(c1, c2small):
    """The clearance-grid broad phase (option clear_grid, SURVEY 8a row a8's AABB pre-reject as an O(1) lookup) is
    conservative: results with the grids on equal the results with them off and the oracle's, and items are dropped."""
    for w, eng, orc in (c1, c2small):
        Q = synth.sample_configs(w.robot, 20000, 23)
        eng.set_option("clear_grid", 0)
        off = eng.feasible_batch(Q)
        eng.set_option("clear_grid", 1)
        eng.set_option("collect_stats", 1); eng.reset_stats()
        on, pairs = eng.feasible_batch(Q, return_pairs=True)
        st = eng.stats()
        eng.set_option("collect_stats", 0); eng.set_option("clear_grid", 0)
        assert np.array_equal(on, off)
        assert st["items_dropped"] > 0
        want = orc.feasible_batch(Q)
        assert_bool_parity(on, want, Q, orc)


def test_edges_floating_base_and_ball_joint(built):
    """Floating / BallAndSocket joints in the edge metric and interpolation (reference
    Cpp/Modeling/Interpolate.cpp:16-52,229-278): same visibility and the same sequential check counts as the oracle."""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_floating()
    eng, orc = Engine(w), OracleWorld(w)
    Q = synth.sample_configs(w.robot, 4000, 31)
    assert_bool_parity(eng.feasible_batch(Q), orc.feasible_batch(Q), Q, orc)
    A, B = synth.sample_edges(w.robot, lambda Q: orc.feasible_batch(Q), 600, 32, rmin=0.1, rmax=1.5)
    vis, nchk = eng.edges_visible_batch(A, B, eps=0.02)
    ovis, onchk = orc.edges_visible_batch(A, B, eps=0.02)
    # midpoints agree to ~1e-15 (device vs host libm), so a disagreement needs a midpoint inside the margin band
    assert (vis != ovis).sum() <= 1 and (nchk != onchk).sum() <= 1
    assert 0.02 < vis.mean() < 0.98
    wts = np.array([2.0, 0.5, 1.5])
    vis, nchk = eng.edges_visible_batch(A[:200], B[:200], eps=0.05, weights=wts)
    ovis, onchk = orc.edges_visible_batch(A[:200], B[:200], eps=0.05, weights=wts)
    assert (vis != ovis).sum() <= 1 and (nchk != onchk).sum() <= 1


def test_multi_link_joint_layout_is_validated_by_the_engine(built):
    from klampt_b200.engine import Engine
    from klampt_b200._capi import KbError
    w = synth.world_floating()
    w.robot.joint_base = None
    with pytest.raises(KbError):
        Engine(w)


def test_solid_box_primitives_pair_queries(built):
    """Box / AABB primitives are solid: explicit geometry-pair queries against the oracle for a box vs a
    mesh (inside / crossing / outside), a point cloud, a sphere, and another box, at random poses."""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(41)
    w = WorldSpec()
    v, t = synth.unit_cube()
    g_small = w.add_geom(GeomSpec.mesh((v - 0.5) * 0.25, t))
    bv, bt = synth.blob_mesh(rng, 2, 0.3)
    g_blob = w.add_geom(GeomSpec.mesh(bv, bt))
    g_cloud = w.add_geom(GeomSpec.cloud(rng.uniform(-0.3, 0.3, size=(400, 3)), rng.uniform(0.0, 0.02, size=400)))
    g_sph = w.add_geom(GeomSpec.sphere([0.05, 0, 0], 0.12))
    g_box = w.add_geom(GeomSpec.box([0.05, -0.1, 0.0], synth.rot_axis_angle([1, 1, 0], 0.6), [0.5, 0.35, 0.25]))
    g_aabb = w.add_geom(GeomSpec.aabb([-0.1, -0.15, -0.2], [0.1, 0.15, 0.2], margin=0.01))
    g_plate = w.add_geom(GeomSpec.aabb([-0.4, -0.3, 0.0], [0.4, 0.3, 0.0]))          # a flat box: only its two faces are surface
    w.robot = synth.make_planar_nR(w, 1)
    eng, orc = Engine(w), OracleWorld(w)
    N = 300
    def poses(spread):
        return np.stack([np.concatenate([synth._random_rotation(rng).reshape(-1), rng.uniform(-spread, spread, size=3)]) for _ in range(N)])
    for ga, gb, spread in ((g_box, g_small, 0.5), (g_small, g_box, 0.5), (g_box, g_blob, 0.6), (g_box, g_cloud, 0.6), (g_sph, g_box, 0.6),
                           (g_box, g_aabb, 0.5), (g_aabb, g_cloud, 0.4), (g_plate, g_blob, 0.5), (g_plate, g_cloud, 0.4)):
        Ta, Tb = poses(spread), poses(spread)
        hit = eng.geom_collides_batch(ga, Ta, gb, Tb)
        d = eng.geom_distance_batch(ga, Ta, gb, Tb)
        near = eng.geom_collides_batch(ga, Ta, gb, Tb, tol=0.05)
        want_hit = np.array([orc.geom_collides(ga, Ta[i], gb, Tb[i]) for i in range(N)])
        want_d = np.array([orc.geom_distance(ga, Ta[i], gb, Tb[i]) for i in range(N)])
        want_near = np.array([orc.geom_within_distance(ga, Ta[i], gb, Tb[i], 0.05) for i in range(N)])
        assert 0.05 < want_hit.mean() < 0.98, (ga, gb, want_hit.mean())
        assert_geom_bool_parity(hit, want_hit, ga, Ta, gb, Tb, orc, 0.0, max_bad=2)
        assert_geom_bool_parity(near, want_near, ga, Ta, gb, Tb, orc, 0.05, max_bad=2)
        np.testing.assert_allclose(d, want_d, rtol=1e-5, atol=1e-9)
    from klampt_b200._capi import KbError
    bad = WorldSpec()
    bad.add_geom(GeomSpec.aabb([0, 0, 0], [1, 0, 0]))                               # two zero dimensions: a segment, not a box
    bad.robot = synth.make_planar_nR(bad, 1)
    with pytest.raises(KbError):
        Engine(bad)


def test_world_with_solid_boxes(built):
    """a robot with box-primitive links among solid boxes, blobs and a solid slab: feasibility, first pairs, clearance and
    edges against the oracle (a link of a chain cannot sit inside a box without a neighbour crossing its surface, so
    containment itself is exercised by test_solid_box_primitives_pair_queries)"""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_boxes()
    eng, orc = Engine(w), OracleWorld(w)
    Q = synth.sample_configs(w.robot, 20000, 43)
    got, pairs = eng.feasible_batch(Q, return_pairs=True)
    want = orc.feasible_batch(Q)
    assert_bool_parity(got, want, Q, orc, max_bad=max(2, len(Q) // 1000))   # distance-threshold elements: in-band mismatches are legal
    assert 0.1 < want.mean() < 0.9
    assert ((pairs[:, 0] >= 0) == (got == 0)).all()
    d = eng.distance_batch(Q[:1500], upper_bound=0.4, include_self=True)
    do, _ = orc.distance_batch(Q[:1500], upper_bound=0.4, include_self=True)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
    A, B = synth.sample_edges(w.robot, lambda X: orc.feasible_batch(X), 300, 44)
    vis, nchk = eng.edges_visible_batch(A, B, eps=0.02)
    ovis, onchk = orc.edges_visible_batch(A, B, eps=0.02)
    assert (vis == ovis).all() and (nchk == onchk).all()
    eng.set_option("clear_grid", 1)
    assert np.array_equal(eng.feasible_batch(Q), got)
    eng.set_option("clear_grid", 0)
    cp, cc = eng.colliding_pairs_batch(Q[:2000], max_pairs=8)
    assert ((cc > 0) == (got[:2000] == 0)).all()


def test_dynamic_point_cloud_rebuilt_on_the_gpu(built):
    """a point cloud obstacle replaced between batches (the reference's dynamic geometries, Cpp/Modeling/ManagedGeometry.h:49-52):
    kb_update_pointcloud rebuilds its hierarchy on the GPU (linear BVH); every state is checked against the oracle built on
    the same points as a static cloud"""
    import copy
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(51)
    base = synth.world_c1()
    T = synth.make_T(synth._random_rotation(rng), [0.1, -0.05, 0.2])
    w = copy.deepcopy(base)
    gdyn = w.add_geom(GeomSpec.dynamic_cloud(200000, radius=0.004, margin=0.002))
    w.objects.append((gdyn, T))
    w.robot = synth.make_arm6(w)                      # ids shift with the extra object: rebuild the robot on the new world
    eng = Engine(w)
    Q = synth.sample_configs(w.robot, 6000, 52)

    def oracle_with(points):
        wo = copy.deepcopy(base)
        if points is not None and len(points):
            wo.objects.append((wo.add_geom(GeomSpec.cloud(points, np.full(len(points), 0.004), margin=0.002)), T))
        else:
            wo.objects.append((-1, T))                # keeps the world ids aligned: an object with empty geometry
        wo.robot = synth.make_arm6(wo)
        return OracleWorld(wo)

    def cloud(n, seed):
        r2 = np.random.default_rng(seed)
        c = r2.uniform([-0.9, -0.9, 0.1], [0.9, 0.9, 1.3], size=(12, 3))
        p = c[r2.integers(0, 12, size=n)] + r2.normal(size=(n, 3)) * 0.05
        return np.ascontiguousarray(p)

    empty = oracle_with(None)
    assert np.array_equal(eng.feasible_batch(Q), empty.feasible_batch(Q))          # before the first update the cloud is empty
    for n, seed in ((150000, 1), (40, 2), (199999, 3), (1, 4), (8, 5)):
        P = cloud(n, seed)
        eng.update_pointcloud(gdyn, P)
        orc = oracle_with(P)
        got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
        assert_bool_parity(got, want, Q, orc, max_bad=max(2, len(Q) // 1000))   # distance-threshold elements: in-band mismatches are legal
        d = eng.distance_batch(Q[:500], upper_bound=0.3)
        do, _ = orc.distance_batch(Q[:500], upper_bound=0.3)
        np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
        if n > 1000:
            assert (got != empty.feasible_batch(Q)).sum() > 0                      # the cloud matters
        # rays see the cloud of the moment too (it is cast through its world-frame group)
        rr = np.random.default_rng(seed)
        src = rr.uniform([-2.5, -2.5, 0.2], [2.5, 2.5, 2.5], (3000, 3))
        rays = np.hstack([src, rr.uniform([-0.9, -0.9, 0.1], [0.9, 0.9, 1.3], (3000, 3)) - src])
        ri, rd, _ = eng.raycast_batch(Q[0], rays)
        oi, od, _ = orc.raycast_batch(Q[0], rays)
        assert np.array_equal(ri, oi)
        np.testing.assert_allclose(rd[ri >= 0], od[oi >= 0], rtol=1e-9, atol=1e-12)
        if n > 1000:
            assert (ri == 11).sum() > 10                                           # the cloud object (world id 11: ground + 10 boxes before it)
    eng.update_pointcloud(gdyn, np.zeros((0, 3)))
    assert np.array_equal(eng.feasible_batch(Q), empty.feasible_batch(Q))
    with pytest.raises(Exception):
        eng.update_pointcloud(gdyn, np.zeros((200001, 3)))


def test_fp32_configurations_entry_point(c1):
    """kb_feasible_batch_f32: float rows are widened on the device; the answer is the fp64 answer for the rounded configuration"""
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 300000, 61)          # large enough for the staged upload
    Qf = Q.astype(np.float32)
    got, pairs = eng.feasible_batch(Qf, return_pairs=True)
    want = eng.feasible_batch(Qf.astype(np.float64))
    assert np.array_equal(got, want)
    assert_bool_parity(got[:5000], orc.feasible_batch(Qf[:5000].astype(np.float64)), Qf[:5000].astype(np.float64), orc)
    assert ((pairs[:, 0] >= 0) == (got == 0)).all()


def test_static_clouds_built_on_the_gpu(c5small):
    """option cloud_builder = 1: the hierarchies of static point clouds come from the GPU builder; same answers"""
    from klampt_b200.engine import Engine
    w, eng, orc = c5small
    eg = Engine(w, options={"cloud_builder": 1})
    Q = synth.sample_configs(w.robot, 8000, 71)
    assert np.array_equal(eg.feasible_batch(Q), eng.feasible_batch(Q))
    d, dg = eng.distance_batch(Q[:800], upper_bound=0.4), eg.distance_batch(Q[:800], upper_bound=0.4)
    np.testing.assert_allclose(dg, d, rtol=1e-12, atol=1e-15)
    # explicit pair queries use the per-geometry hierarchy, also built on the GPU
    gi = [i for i, g in enumerate(w.geoms) if g.kind == "cloud"][0]
    gl = w.robot.link_geom[4]
    rng = np.random.default_rng(72)
    Ta = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (200, 1))
    Tb = np.stack([np.concatenate([synth._random_rotation(rng).reshape(-1), rng.uniform(-1, 1, size=3) + [0, 0, 0.8]]) for _ in range(200)])
    assert np.array_equal(eg.geom_collides_batch(gi, Ta, gl, Tb), eng.geom_collides_batch(gi, Ta, gl, Tb))


def test_environment_mesh_built_on_the_gpu(c2small):
    """option mesh_builder = 1: the merged environment mesh hierarchy comes from the GPU builder (Morton order of the triangle
    centroids, one triangle per leaf); same answers as with the host SAH tree"""
    from klampt_b200.engine import Engine
    w, eng, orc = c2small
    eg = Engine(w, options={"mesh_builder": 1})
    assert eg.layout()["nodes"] > eng.layout()["nodes"]          # the reserved node range of the GPU-built tree (2 n + 4)
    Q = synth.sample_configs(w.robot, 20000, 81)
    got, pairs = eg.feasible_batch(Q, return_pairs=True)
    assert np.array_equal(got, eng.feasible_batch(Q))
    assert_bool_parity(got[:4000], orc.feasible_batch(Q[:4000]), Q[:4000], orc)
    np.testing.assert_allclose(eg.distance_batch(Q[:1000], upper_bound=0.3, include_self=True), eng.distance_batch(Q[:1000], upper_bound=0.3, include_self=True),
                               rtol=1e-12, atol=1e-15)


def test_degenerate_triangles_agree_with_the_oracle(built):
    """meshes in the wild carry zero-area triangles (repeated or collinear vertices): whatever the predicates make of them, the
    engine and the oracle must make the same thing"""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(97)
    w = synth.world_c1()
    extra_v, extra_t = [], []
    for k in range(60):
        a = rng.uniform([-0.8, -0.8, 0.1], [0.8, 0.8, 1.2]); d = rng.normal(size=3) * 0.15
        b = a + d
        c = b if k % 2 == 0 else a + 0.37 * d             # repeated vertex / three collinear vertices
        base = 3 * k
        extra_v += [a, b, c]; extra_t.append([base, base + 1, base + 2])
    w.objects.append((w.add_geom(GeomSpec.mesh(np.array(extra_v), np.array(extra_t, dtype=np.int32))), synth.make_T(None, (0, 0, 0))))
    w.robot = synth.make_arm6(w)
    eng, orc = Engine(w), OracleWorld(w)
    Q = synth.sample_configs(w.robot, 20000, 98)
    got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
    base = OracleWorld(synth.world_c1()).feasible_batch(Q)
    assert (want != base).sum() > 0                        # the slivers do get hit
    assert_bool_parity(got, want, Q, orc, max_bad=max(2, len(Q) // 1000))   # distance-threshold elements: in-band mismatches are legal
    d = eng.distance_batch(Q[:1000], upper_bound=0.3)
    do, _ = orc.distance_batch(Q[:1000], upper_bound=0.3)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_segment_primitives_agree_with_the_oracle(built):
    """GeometricPrimitive "Segment": bars with a margin (capsules) around the arm, and segment-vs-{segment, sphere, box, mesh} pair
    queries; the margin sends every pair that involves the zero-area element through the fp64 path"""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    rng = np.random.default_rng(131)
    w = synth.world_c1()
    for k in range(30):
        a = rng.uniform([-0.8, -0.8, 0.1], [0.8, 0.8, 1.2]); b = a + rng.normal(size=3) * 0.2
        w.objects.append((w.add_geom(GeomSpec.segment(a, b, margin=0.0 if k % 3 == 0 else 0.01)), synth.make_T(synth._random_rotation(rng), rng.uniform(-0.1, 0.1, 3))))
    gseg = [w.add_geom(GeomSpec.segment(rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3), margin=m)) for m in (0.0, 0.0, 0.02)]
    gsph = w.add_geom(GeomSpec.sphere([0.05, 0, 0], 0.1))
    gbox = w.add_geom(GeomSpec.box([0, 0, 0], synth._random_rotation(rng), [0.2, 0.1, 0.05]))
    w.robot = synth.make_arm6(w)
    glink = w.robot.link_geom[3]
    eng, orc = Engine(w), OracleWorld(w)
    Q = synth.sample_configs(w.robot, 20000, 132)
    got, want = eng.feasible_batch(Q), orc.feasible_batch(Q)
    assert (want != OracleWorld(synth.world_c1()).feasible_batch(Q)).sum() > 0          # the bars do get hit
    assert_bool_parity(got, want, Q, orc, max_bad=max(2, len(Q) // 1000))   # distance-threshold elements: in-band mismatches are legal
    d = eng.distance_batch(Q[:1000], upper_bound=0.3)
    do, _ = orc.distance_batch(Q[:1000], upper_bound=0.3)
    np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-9)
    n = 400
    Ta = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-0.25, 0.25, 3)) for _ in range(n)])
    Tb = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-0.25, 0.25, 3)) for _ in range(n)])
    for ga, gb in ((gseg[0], gseg[1]), (gseg[0], gseg[2]), (gseg[2], gsph), (gseg[1], gbox), (gseg[2], gbox), (gseg[0], glink), (glink, gseg[2])):
        dg = eng.geom_distance_batch(ga, Ta, gb, Tb)
        dw = np.array([orc.geom_distance(ga, Ta[i], gb, Tb[i]) for i in range(n)])
        np.testing.assert_allclose(dg, dw, rtol=1e-5, atol=1e-9)
        for tol in (0.0, 0.03):
            cg = eng.geom_collides_batch(ga, Ta, gb, Tb, tol=tol).astype(bool)
            cw = np.array([orc.geom_within_distance(ga, Ta[i], gb, Tb[i], tol) if tol > 0 else orc.geom_collides(ga, Ta[i], gb, Tb[i]) for i in range(n)])
            near = np.abs(dw - tol) < 1e-6                                              # the stated band of the boolean parity
            assert np.array_equal(cg[~near], cw[~near])


def test_closest_points_element_ids_and_tolerances(built):
    """DistanceQueryResult beyond the scalar (src/geometry.h:631-694): closest points on the margin-inflated surfaces, element
    indices in the geometries' own order, and the absErr / relErr contract of AnyCollisionQuery::Distance.  The points are not
    compared with anything the oracle stores: they are RE-MEASURED by the oracle -- |cp2 - cp1| must be d, and each point must lie on
    its geometry (oracle point-to-geometry distance = 0)."""
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld, point_tri_distance
    rng = np.random.default_rng(11)
    w = WorldSpec()
    vb, tb = synth.blob_mesh(rng, 2, 0.4)
    vc, tc = synth.unit_cube()
    ga = w.add_geom(GeomSpec.mesh(vb, tb))
    gb = w.add_geom(GeomSpec.mesh(vc, tc))
    gm = w.add_geom(GeomSpec.mesh(vc, tc, margin=0.03))
    pts = rng.uniform(-0.3, 0.3, size=(3000, 3))
    gc = w.add_geom(GeomSpec.cloud(pts, None, margin=0.01))
    gs = w.add_geom(GeomSpec.sphere([0.1, 0.0, 0.0], 0.2))
    gp = w.add_geom(GeomSpec.point([0.0, 0.0, 0.0]))
    w.robot = synth.make_planar_nR(w, 2)
    eng, orc = Engine(w), OracleWorld(w)
    N = 300
    Ta = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-0.2, 0.2, size=3)) for _ in range(N)])
    Tb = np.stack([synth.make_T(synth._random_rotation(rng), rng.uniform(-1.8, 1.8, size=3)) for _ in range(N)])
    margin = {ga: 0.0, gb: 0.0, gm: 0.03, gc: 0.01, gs: 0.0}

    def at(p):
        T = synth.IDENTITY12.copy(); T[9:12] = p
        return T
    for g1, g2 in ((ga, gb), (ga, gm), (gb, gc), (gs, ga), (gc, gs), (gm, gc)):
        d, cp, el = eng.geom_distance_batch_ex(g1, Ta, g2, Tb)
        do = np.array([orc.geom_distance(g1, Ta[i], g2, Tb[i]) for i in range(N)])
        np.testing.assert_allclose(d, do, rtol=1e-5, atol=1e-12)
        far = d > 1e-9
        assert far.sum() > 50
        np.testing.assert_allclose(np.linalg.norm(cp[far, 1] - cp[far, 0], axis=1), d[far], rtol=1e-9, atol=1e-12)
        for i in np.nonzero(far)[0][:60]:
            # the oracle's distance from a point to the geometry subtracts the geometry's margin: 0 on the inflated surface
            assert abs(orc.geom_distance(g1, Ta[i], gp, at(cp[i, 0]))) < 1e-9, (g1, g2, i)
            assert abs(orc.geom_distance(g2, Tb[i], gp, at(cp[i, 1]))) < 1e-9, (g1, g2, i)
        for i in np.nonzero(~far & (d == 0))[0][:40]:         # intersecting meshes: one common point, on both surfaces
            assert np.array_equal(cp[i, 0], cp[i, 1])
            assert abs(orc.geom_distance(g1, Ta[i], gp, at(cp[i, 0])) + margin[g1]) < 1e-9
            assert abs(orc.geom_distance(g2, Tb[i], gp, at(cp[i, 1])) + margin[g2]) < 1e-9
    # element indices refer to the caller's arrays: the reported triangle of the blob really is where cp1 lies
    d, cp, el = eng.geom_distance_batch_ex(ga, Ta, gb, Tb)
    for i in np.nonzero(d > 1e-9)[0][:80]:
        tri = synth.transform_points(Ta[i], vb[tb[el[i, 0]]])
        assert point_tri_distance(cp[i, 0], tri) < 1e-9
        tri2 = synth.transform_points(Tb[i], vc[tc[el[i, 1]]])
        assert point_tri_distance(cp[i, 1], tri2) < 1e-9
    d, cp, el = eng.geom_distance_batch_ex(gb, Ta, gc, Tb)
    for i in np.nonzero(d > 1e-9)[0][:80]:
        p = synth.transform_points(Tb[i], pts[el[i, 1]][None])[0]
        assert abs(np.linalg.norm(p - cp[i, 1]) - 0.01) < 1e-9          # the cloud's point, moved by the cloud's margin towards the cube
    # tolerances: never below the exact value, never more than the tolerance above it; and cheaper
    exact = eng.geom_distance_batch(ga, Ta, gb, Tb)
    for ae, re_ in ((0.05, 0.0), (0.0, 0.2)):
        dt, _, _ = eng.geom_distance_batch_ex(ga, Ta, gb, Tb, abs_err=ae, rel_err=re_)
        assert (dt >= exact - 1e-12).all() and (dt <= exact + ae + re_ * np.abs(exact) + 1e-9).all()
    # the robot query: same distances as the plain entry point, points on the reported link / obstacle
    wc = synth.world_c1()
    gp2 = wc.add_geom(GeomSpec.point([0.0, 0.0, 0.0]))
    e2, o2 = Engine(wc), OracleWorld(wc)
    Q = synth.sample_configs(wc.robot, 400, 5)
    d, pairs, cp, el = e2.distance_batch_ex(Q, upper_bound=0.5, include_self=True)
    d0, p0 = e2.distance_batch(Q, upper_bound=0.5, include_self=True, return_pairs=True)
    assert np.array_equal(d, d0) and np.array_equal(pairs, p0)
    nid = e2.num_ids(); L = wc.robot.L; base = nid - L
    T = o2.fk_batch(Q)
    seen = 0
    for i in range(len(Q)):
        if not (1e-9 < d[i] < 0.5):
            assert d[i] <= 1e-9 or np.isnan(cp[i]).all()
            continue
        assert abs(np.linalg.norm(cp[i, 1] - cp[i, 0]) - d[i]) < 1e-9
        a = pairs[i, 0]
        assert a >= base                                          # first id: a robot link (CheckCollisionFree order)
        assert abs(o2.geom_distance(wc.robot.link_geom[a - base], T[i, a - base], gp2, at(cp[i, 0]))) < 1e-9
        seen += 1
    assert seen > 100


def test_edges_with_non_finite_or_absurd_length(c1):
    """an edge with a NaN / inf coordinate is invisible without a single check; an edge that would need more than 2^24 pieces is an
    error, never a coarser check than the caller asked for"""
    from klampt_b200._capi import KbError
    w, eng, orc = c1
    Q = synth.sample_configs(w.robot, 64, 90)
    ok = Q[orc.feasible_batch(Q) == 1]
    A, B = ok[:8].copy(), ok[8:16].copy()
    B[1, 2] = np.inf; A[3, 1] = np.nan
    vis, n = eng.edges_visible_batch(A, B, eps=0.05)
    ovis, on = orc.edges_visible_batch(np.delete(A, [1, 3], 0), np.delete(B, [1, 3], 0), eps=0.05)
    assert vis[1] == 0 and vis[3] == 0 and n[1] == 0 and n[3] == 0
    assert np.array_equal(np.delete(vis, [1, 3]), ovis) and np.array_equal(np.delete(n, [1, 3]), on)
    with pytest.raises(KbError):
        eng.edges_visible_batch(ok[:2], ok[2:4], eps=1e-9)


def test_small_edge_batches_all_levels_at_once(c1, c2small):
    """a small edge batch is checked with every midpoint of every level in one launch (no early exit, no per-level read-back); visibility
    and the sequential checker's nchecks must equal the level-by-level form and the oracle"""
    for (w, eng, orc), eps in ((c1, 0.01), (c2small, 0.02)):
        A, B = synth.sample_edges(w.robot, lambda Q: eng.feasible_batch(Q), 700, 17)
        for n in (1, 5, 64, 700):
            vis, nchk = eng.edges_visible_batch(A[:n], B[:n], eps=eps)
            ovis, on = orc.edges_visible_batch(A[:n], B[:n], eps=eps)
            assert np.array_equal(vis, ovis) and np.array_equal(nchk, on), n
        eng.set_option("edge_flat_max", 0)
        try:
            v2, n2 = eng.edges_visible_batch(A, B, eps=eps)
        finally:
            eng.set_option("edge_flat_max", 1 << 18)
        v1, n1 = eng.edges_visible_batch(A, B, eps=eps)
        assert np.array_equal(v1, v2) and np.array_equal(n1, n2)
        # degenerate edges: a == b needs no check at all
        vis, nchk = eng.edges_visible_batch(A[:3], A[:3], eps=eps)
        assert vis.all() and (nchk == 0).all()
