"""Host logic of the reference-facing adapters that does not need a GPU: WorldCollider mask semantics against the
oracle's InitializeDefault mask, CSpace bookkeeping, world <-> spec round trip, sharding arithmetic."""
import numpy as np
import pytest

from klampt_b200 import synth
from klampt_b200.collide import WorldCollider, bb_intersect
from klampt_b200.cspace import CSpace
from klampt_b200.robotsim import WorldModel, Geometry3D, TriangleMesh
from klampt_b200.shard import shard_range, interleaved_indices
from oracle.oracle import OracleWorld


def test_world_spec_round_trip():
    spec = synth.world_c1()
    world = WorldModel.from_spec(spec)
    assert world.numTerrains() == 1 and world.numRigidObjects() == 10 and world.numRobots() == 1
    assert world.numIDs() == spec.num_ids() and world.robotLinkID(0, 3) == spec.robot_link_id(3)
    back = world.to_spec()
    assert back.total_tris() == spec.total_tris()
    for (g0, T0), (g1, T1) in zip(spec.objects, back.objects):
        assert np.allclose(T0, T1)
        assert np.array_equal(spec.geoms[g0].tris, back.geoms[g1].tris)
    r = world.robot(0)
    assert r.numLinks() == 7 and r.link(2).getParent() == 1 and not r.link(2).isPrismatic()
    R, t = r.link(3).getParentTransform()
    assert np.allclose(t, spec.robot.T0[3, 9:]) and len(R) == 9
    assert r.selfCollisionEnabled(0, 2) and not r.selfCollisionEnabled(2, 3) and r.selfCollisionEnabled(3, 1)
    r.enableSelfCollision(0, 2, False)
    assert not r.selfCollisionEnabled(2, 0)


def test_world_collider_mask_matches_initialize_default():
    """collide.py:326-359 restates PlannerSettings.cpp:16-41; both must enable exactly the same robot pairs"""
    spec = synth.world_c1()
    world = WorldModel.from_spec(spec)
    col = WorldCollider(world)
    m = col.to_pair_mask()
    ref = OracleWorld(spec).pair_mask()
    rid, base, L = spec.robot_id(), spec.robot_link_id(0), spec.robot.L
    for j in range(L):
        for s in range(rid):                                   # link vs terrain / object
            assert bool(m[base + j, s] or m[s, base + j]) == bool(ref[base + j, s] or ref[s, base + j])
        for k in range(j + 1, L):                              # self pairs, upper triangular
            assert m[base + j, base + k] == ref[base + j, base + k]
    # ignoreCollision of one pair and of a whole body
    col.ignoreCollision((world.robot(0).link(6), world.rigidObject(0)))
    col.ignoreCollision(world.rigidObject(1))
    m2 = col.to_pair_mask()
    assert m2[base + 6, spec.rigid_object_id(0)] == 0 and m2[spec.rigid_object_id(0), base + 6] == 0
    assert not m2[:, spec.rigid_object_id(1)].any() and not m2[spec.rigid_object_id(1), :].any()
    assert m2[base + 5, spec.rigid_object_id(0)] == 1
    assert not col.isCollisionEnabled(world.rigidObject(1)) and col.isCollisionEnabled((world.robot(0).link(5), world.rigidObject(0)))


def test_geometry_bb_is_loose_but_conservative():
    v, t = synth.unit_cube()
    g = Geometry3D(TriangleMesh(v, t))
    R = synth.rot_axis_angle([0, 0, 1], 0.7)
    g.setCurrentTransform(list(R.T.reshape(-1)), [1, 2, 3])
    lo, hi = g.getBB()
    lt, ht = g.getBBTight()
    assert all(a <= b + 1e-12 for a, b in zip(lo, lt)) and all(a >= b - 1e-12 for a, b in zip(hi, ht))
    g.setCollisionMargin(0.1)
    lo2, _ = g.getBB()
    assert np.allclose(np.array(lo) - np.array(lo2), 0.1)
    assert bb_intersect((lo, hi), (lt, ht)) and not bb_intersect((lo, hi), ([9, 9, 9], [10, 10, 10]))


def test_cspace_bookkeeping_and_edge_checker():
    class Disk(CSpace):
        def __init__(self):
            CSpace.__init__(self)
            self.setBounds([(0, 1), (0, 1)])
            self.eps = 0.01
            self.addFeasibilityTest(lambda q: self.inBounds(q), "bounds")
            self.addFeasibilityTest(lambda q: (q[0] - 0.5) ** 2 + (q[1] - 0.5) ** 2 > 0.09, "disk", dependencies="bounds")

    s = Disk()
    s.setup()
    assert s.properties["volume"] == 1 and s.feasibilityTestDependencies == [("disk", "bounds")]
    assert s.isFeasible([0.1, 0.1]) and not s.isFeasible([0.5, 0.5]) and not s.isFeasible([1.5, 0.5])
    assert s.feasibilityFailures([0.5, 0.5]) == ["disk"]
    assert s.isVisible([0.1, 0.1], [0.9, 0.1]) and not s.isVisible([0.1, 0.5], [0.9, 0.5])
    st = s.getStats()
    assert st["feasible_count"] == 3 and st["visible_count"] == 2 and st["visible_probability"] == 0.5
    assert st["disk_count"] == 2 and st["bounds_count"] == 3 and st["average_visible_length"] == pytest.approx(0.8)
    s.close()
    assert s.cspace is None


def test_shard_arithmetic():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert sum(hi - lo for lo, hi in blocks) == n
            idx = np.concatenate([interleaved_indices(n, r, world, 16) for r in range(world)])
            assert sorted(idx) == list(range(n))
    with pytest.raises(ValueError):
        shard_range(10, 3, 2)


def test_robot_model_metric_and_interpolation_match_the_oracle_for_multi_link_joints():
    """RobotModel.interpolate / distance (the host-side scalar forms of Klampt::Interpolate / Distance,
    Cpp/Modeling/Interpolate.cpp:10-71,208-343) agree with the oracle for Floating and BallAndSocket joints."""
    spec = synth.world_floating()
    world = WorldModel.from_spec(spec)
    r = world.robot(0)
    orc = OracleWorld(spec)
    Q = synth.sample_configs(spec.robot, 80, 9)
    rng = np.random.default_rng(2)
    for i in range(40):
        a, b, u = Q[2 * i], Q[2 * i + 1], rng.uniform()
        assert abs(r.distance(a, b) - orc.cspace_distance(a, b)) < 1e-9
        np.testing.assert_allclose(r.interpolate(a, b, u), orc.interpolate(a, b, u), atol=1e-9)
    back = world.to_spec()
    assert np.array_equal(back.robot.joint_base, spec.robot.joint_base)


def test_cspace_named_test_queries():
    """CSpaceInterface's per-test queries (Python/klampt/src/motionplanning.h:122-171) on the host-side CSpace mirror"""
    sp = CSpace()
    sp.eps = 0.05
    sp.setBounds([(0.0, 1.0), (0.0, 1.0)])
    sp.addFeasibilityTest(lambda q: sp.inBounds(q), "bounds")
    sp.addFeasibilityTest(lambda q: (q[0] - 0.5) ** 2 + (q[1] - 0.5) ** 2 > 0.04, "disk", dependencies="bounds")
    sp.addFeasibilityTest(lambda q: q[1] < 0.9, "ceiling")
    assert sp.feasibilityQueryOrder() == ["bounds", "disk", "ceiling"] and sp.feasibilityTestDependenciesOf("disk") == ["bounds"]
    assert sp.testFeasibility("disk", [0.5, 0.5]) is False and sp.testFeasibility("ceiling", [0.5, 0.5]) is True
    assert sp.feasibilityFailures([0.5, 0.95]) == ["ceiling"] and sp.feasibilityFailures([0.5, 0.5]) == ["disk"]
    with pytest.raises(ValueError):
        sp.testFeasibility("nope", [0, 0])
    a, b = [0.1, 0.5], [0.9, 0.5]
    assert not sp.testVisibility("disk", a, b) and sp.testVisibility("ceiling", a, b)
    assert sp.visibilityFailures(a, b) == ["disk"] and sp.visibilityFailures([0.1, 0.1], [0.9, 0.1]) == []
    assert not sp.isVisible(a, b) and sp.isVisible([0.1, 0.1], [0.9, 0.1])
    with pytest.raises(ValueError):
        sp.setVisibilityEpsilon(0.0)
    sp.setVisibilityEpsilon(0.01)
    assert sp.eps == 0.01


def test_adaptive_queries_reorder_the_conjunction():
    """CSpaceInterface's adaptive-query face (motionplanning.h:139-166): per-test running cost / probability, priors, dependencies,
    and an optimised order that runs cheap, selective tests first"""
    from klampt_b200.cspace import CSpace
    calls = []
    sp = CSpace()
    sp.bound = [(0.0, 1.0)]
    sp.addFeasibilityTest(lambda x: calls.append("setup") or True, "setup")
    sp.addFeasibilityTest(lambda x: calls.append("slow") or True, "slow", dependencies="setup")
    sp.addFeasibilityTest(lambda x: calls.append("picky") or x[0] < 0.2, "picky")
    with pytest.raises(RuntimeError, match="adaptive queries not enabled"):
        sp.optimizeQueryOrder()
    assert not sp.adaptiveQueriesEnabled() and sp.feasibilityQueryOrder() == ["setup", "slow", "picky"]
    sp.enableAdaptiveQueries()
    assert sp.adaptiveQueriesEnabled()
    sp.setFeasibilityPrior("setup", 1.0, 0.99, 10.0)
    sp.setFeasibilityPrior("slow", 50.0, 0.9, 10.0)
    sp.setFeasibilityPrior("picky", 1.0, 0.2, 10.0)
    assert sp.feasibilityCost("slow") == 50.0 and sp.feasibilityProbability("picky") == 0.2
    sp.optimizeQueryOrder()
    assert sp.feasibilityQueryOrder() == ["picky", "setup", "slow"]           # 1/0.8 < 1/0.01 < 50/0.1, and slow still follows setup
    calls.clear()
    assert not sp.isFeasible([0.7]) and calls == ["picky"]                    # the selective test fails first: nothing else runs
    calls.clear()
    assert sp.isFeasible([0.1]) and calls == ["picky", "setup", "slow"]
    assert 0.2 < sp.feasibilityProbability("picky") < 0.3                     # (0.2 * 10 + 0 + 1) / 12
    sp.setFeasibilityDependency("picky", "slow")                              # now picky may only run after slow
    sp.optimizeQueryOrder()
    assert sp.feasibilityQueryOrder() == ["setup", "slow", "picky"]
    with pytest.raises(ValueError, match="Invalid dependency"):
        sp.setFeasibilityDependency("picky", "nope")
    with pytest.raises(ValueError, match="Invalid constraint name"):
        sp.feasibilityCost("nope")
    sp.setFeasibilityDependency("setup", "picky")                             # closes a cycle
    with pytest.raises(ValueError, match="Invalid dependency"):
        sp.optimizeQueryOrder()
    sp.setVisibilityPrior("slow", 3.0, 0.5, 1.0)
    assert sp.visibilityCost("slow") == 3.0 and sp.visibilityProbability("slow") == 0.5 and len(sp.visibilityQueryOrder()) == 3


def test_other_robots_ride_along_as_rigid_bodies():
    """SingleRobotCSpace::CheckCollisionFree checks the robot against all OTHER robots too (RobotCSpace.cpp:794-823): WorldModel.to_spec
    hands their links to the engine as rigid objects at their current transforms and keeps the id translation"""
    from klampt_b200 import so3
    spec = synth.world_c1()
    world = WorldModel.from_spec(spec)
    other = world.addRobot("second", spec.robot, spec.geoms)
    T_other = synth.make_T(None, (0.9, 0.25, 0.0))
    for j in range(other.numLinks()):                     # (setConfig needs the GPU: place the second arm's links by hand)
        other.link(j).geometry().setCurrentTransform(*so3.from_rowmajor12(T_other))
    T, O, L = world.numTerrains(), world.numRigidObjects(), spec.robot.L
    n_geo = sum(1 for j in range(L) if not other.link(j).geometry().empty())
    w0 = world.to_spec(0)
    assert len(w0.terrains) == T and len(w0.objects) == O + L
    assert sum(1 for gi, _ in w0.objects[O:] if gi >= 0) == n_geo
    assert all(np.allclose(Tm, T_other) for _, Tm in w0.objects[O:])
    ids = list(w0.world_ids)
    assert ids[:T + O] == list(range(T + O))
    assert ids[T + O:T + O + L] == [world.robotLinkID(1, j) for j in range(L)]
    assert ids[T + O + L] == world.robotID(0) and ids[T + O + L + 1:] == [world.robotLinkID(0, j) for j in range(L)]
    # the second robot as the active one: the first one's links ride along
    w1 = world.to_spec(1)
    assert list(w1.world_ids[T + O:T + O + L]) == [world.robotLinkID(0, j) for j in range(L)] and w1.world_ids[T + O + L] == world.robotID(1)
    # a reference-numbered mask is permuted into the engine's numbering
    wc = WorldCollider(world)
    m = wc.to_pair_mask()
    w0m = world.to_spec(0, pair_mask=m)
    a, b = world.robotLinkID(0, 3), world.robotLinkID(1, 2)
    ea, eb = ids.index(a), ids.index(b)
    assert m[min(a, b), max(a, b)] == 1 and (w0m.pair_mask[ea, eb] == 1 or w0m.pair_mask[eb, ea] == 1)
    assert w0m.pair_mask.shape == (len(ids), len(ids))
    # the oracle agrees that a link of the first arm placed inside the second arm collides with it (and names an object id)
    single = WorldModel.from_spec(spec).to_spec(0)
    assert single.world_ids is None and len(single.objects) == O
    Q = synth.sample_configs(spec.robot, 3000, 5)
    f_two, f_one = OracleWorld(world.to_spec(0)).feasible_batch(Q), OracleWorld(single).feasible_batch(Q)
    assert np.all(f_two <= f_one) and (f_two < f_one).sum() > 0          # the second arm only takes feasible configurations away
