"""The reference-facing Python adapters (RobotCSpace / WorldCollider / robotsim mirror) on the GPU, checked against the
oracle; these read like the calls a Klamp't user makes (plan/robotcspace.py, model/collide.py, robotsim)."""
import numpy as np
import pytest

from klampt_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(built):
    from klampt_b200.collide import WorldCollider
    from klampt_b200.robotcspace import RobotCSpace
    from klampt_b200.robotsim import WorldModel
    from oracle.oracle import OracleWorld
    spec = synth.world_c1()
    world = WorldModel.from_spec(spec)
    collider = WorldCollider(world)
    space = RobotCSpace(world.robot(0), collider)
    return spec, world, collider, space, OracleWorld(spec)


def test_robot_cspace_batch_and_single(setup):
    spec, world, collider, space, orc = setup
    Q = synth.sample_configs(spec.robot, 3000, 21)
    want = orc.feasible_batch(Q)
    assert np.array_equal(space.feasible_batch(Q), want)
    space.setup()
    for i in range(25):
        assert space.feasible(list(Q[i])) == bool(want[i])
        assert space.isFeasible(list(Q[i])) == bool(want[i])
    st = space.getStats()
    assert st["feasible_count"] == 25 and st["engine_configs_checked"] >= 3000
    assert len(space.bound) == 7 and space.properties["geodesic"] == 1 and space.eps == 1e-2


def test_robot_cspace_visibility(setup):
    spec, world, collider, space, orc = setup
    A, B = synth.sample_edges(spec.robot, lambda X: orc.feasible_batch(X), 400, 22)
    vis, n = space.visible_batch(A, B, return_nchecks=True)
    ovis, on = orc.edges_visible_batch(A, B, eps=space.eps)
    assert np.array_equal(vis, ovis) and np.array_equal(n, on)
    for i in range(10):
        assert space.isVisible(list(A[i]), list(B[i])) == bool(ovis[i])
        assert space.distance(A[i], B[i]) == pytest.approx(orc.cspace_distance(A[i], B[i]))
        assert np.allclose(space.interpolate(A[i], B[i], 0.25), orc.interpolate(A[i], B[i], 0.25))


def test_ignore_collision_reaches_the_kernels(setup):
    from klampt_b200.collide import WorldCollider
    from klampt_b200.robotcspace import RobotCSpace
    from klampt_b200.robotsim import WorldModel
    from oracle.oracle import OracleWorld
    spec, world, collider, space, orc = setup
    w2 = WorldModel.from_spec(spec)
    col = WorldCollider(w2, ignore=[w2.rigidObject(i) for i in range(w2.numRigidObjects())])
    s2 = RobotCSpace(w2.robot(0), col)
    spec2 = synth.world_c1()
    spec2.objects = []                      # same world without the rigid objects
    spec2.pair_mask = None
    Q = synth.sample_configs(spec.robot, 3000, 23)
    o2 = OracleWorld(spec2)
    assert np.array_equal(s2.feasible_batch(Q), o2.feasible_batch(Q))
    assert s2.feasible_batch(Q).sum() > space.feasible_batch(Q).sum()


def test_robotsim_setconfig_selfcollides_and_geometry_queries(setup):
    spec, world, collider, space, orc = setup
    robot = world.robot(0)
    Q = synth.sample_configs(spec.robot, 400, 24)
    # self-collision only oracle: same robot, no environment
    from klampt_b200.worldspec import WorldSpec
    from oracle.oracle import OracleWorld
    ws = WorldSpec()
    ws.robot = synth.make_arm6(ws)
    ws.robot.qmin[:] = -np.inf
    ws.robot.qmax[:] = np.inf
    o_self = OracleWorld(ws)
    want = o_self.feasible_batch(Q) == 0
    assert np.array_equal(robot.selfCollidesBatch(Q), want)
    for i in list(np.nonzero(want)[0][:3]) + list(np.nonzero(~want)[0][:3]):
        robot.setConfig(list(Q[i]))
        assert robot.selfCollides() == bool(want[i])
        assert robot.getConfig() == list(Q[i])
        T = orc.fk(Q[i])
        R, t = robot.link(6).getTransform()
        assert np.allclose(np.array(R).reshape(3, 3).T, T[6, :9].reshape(3, 3), atol=1e-12) and np.allclose(t, T[6, 9:], atol=1e-12)
        # per-pair Geometry3D queries at the configuration just set (robotsim.cpp:1656-1819)
        g6, gobj = robot.link(6).geometry(), world.rigidObject(0).geometry()
        gi_link, gi_obj = spec.robot.link_geom[6], spec.objects[0][0]
        d_want = orc.geom_distance(gi_link, T[6], gi_obj, spec.objects[0][1])
        assert g6.distance(gobj).d == pytest.approx(d_want, rel=1e-5, abs=1e-12)
        assert g6.collides(gobj) == (d_want <= 0)
        assert g6.withinDistance(gobj, d_want + 1e-3) and (d_want <= 1e-3 or not g6.withinDistance(gobj, d_want - 1e-3))
    with pytest.raises(ValueError):
        robot.setConfig([0.0] * 3)
    hits = list(collider.robotObjectCollisions(0))
    robot.setConfig(list(Q[0]))
    assert isinstance(hits, list)
