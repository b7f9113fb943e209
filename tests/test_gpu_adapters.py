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
    assert len(space.bound) == 7 and space.properties["geodesic"] == 1 and space.eps == 1e-3


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


def test_cpp_batch_single_robot_cspace(built, tmp_path):
    """include/klampt_b200/BatchSingleRobotCSpace.h (the C++ face of SingleRobotCSpace) compiled with g++ against the C ABI"""
    import os
    import shutil
    import struct
    import subprocess
    from oracle.oracle import OracleWorld
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "test_batch_cspace")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "test_batch_cspace.cpp"),
                           "-o", exe, "-L", os.path.join(root, "klampt_b200"), "-lklampt_b200", "-Wl,-rpath," + os.path.join(root, "klampt_b200")])
    spec = synth.world_c1()
    orc = OracleWorld(spec)
    Q = synth.sample_configs(spec.robot, 2000, 31)
    A, B = synth.sample_edges(spec.robot, lambda X: orc.feasible_batch(X), 200, 32)

    def blob(a, dt):
        a = np.ascontiguousarray(a, dtype=dt).reshape(-1)
        return struct.pack("<q", a.size) + a.tobytes()

    r = spec.robot
    with open(tmp_path / "world.bin", "wb") as f:
        f.write(blob([len(spec.geoms)], np.int64))
        for g in spec.geoms:
            f.write(blob(g.verts, np.float64) + blob(g.tris, np.int32))
        f.write(blob(spec.terrains, np.int32))
        f.write(blob([g for g, _ in spec.objects], np.int32) + blob(np.array([T for _, T in spec.objects]), np.float64))
        f.write(blob(r.parents, np.int32) + blob(r.linktype, np.uint8) + blob(r.axis, np.float64) + blob(r.T0, np.float64) + blob(r.qmin, np.float64)
                + blob(r.qmax, np.float64) + blob(r.link_geom, np.int32) + blob(r.joint_type, np.uint8) + blob(r.joint_link, np.int32))
        f.write(blob(Q, np.float64) + blob(A, np.float64) + blob(B, np.float64))
    res = subprocess.run([exe, str(tmp_path / "world.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    feas = np.frombuffer(raw[:2000], dtype=np.uint8)
    vis = np.frombuffer(raw[2000:2200], dtype=np.uint8)
    nchk = np.frombuffer(raw[2200:3000], dtype=np.int32)
    assert np.array_equal(feas, orc.feasible_batch(Q))
    ovis, on = orc.edges_visible_batch(A, B, eps=0.01)
    assert np.array_equal(vis, ovis) and np.array_equal(nchk, on)
    # the ray casts of the C++ face (RayCastBatch): 400 rays it built itself, robot at Q[0]
    rays = np.frombuffer(raw[3000:3000 + 400 * 48], dtype=np.float64).reshape(400, 6)
    rid = np.frombuffer(raw[3000 + 400 * 48:3000 + 400 * 52], dtype=np.int32)
    rdist = np.frombuffer(raw[3000 + 400 * 52:], dtype=np.float64)
    oid, od, _ = orc.raycast_batch(Q[0], rays)
    assert np.array_equal(rid, oid) and (rid >= 0).sum() > 100
    np.testing.assert_allclose(rdist[rid >= 0], od[oid >= 0], rtol=1e-9, atol=1e-12)


def test_batch_roadmap_planner_on_the_engine(setup):
    """klampt_b200.plan.MotionPlan (PRM / Lazy-PRM* / RRT / SBL over feasible_batch + visible_batch): the returned path is verified
    milestone by milestone and edge by edge with the CPU oracle"""
    from klampt_b200.plan import MotionPlan
    spec, world, collider, space, orc = setup
    Q = synth.sample_configs(spec.robot, 400, 61)
    feas = Q[orc.feasible_batch(Q) == 1]
    start, goal = feas[0], feas[7]
    for kind in ("prm", "lazyprm*", "rrt", "sbl"):
        tree = kind in ("rrt", "sbl")
        plan = MotionPlan(space, kind, knn=10, batch=256 if tree else 1500, seed=5, perturbationRadius=0.5)
        plan.setEndpoints(list(start), list(goal))
        path = None
        for _ in range(60 if tree else 6):
            plan.planMore(1)
            path = plan.getPath()
            if path:
                break
        assert path is not None, kind
        P = np.array(path)
        assert np.allclose(P[0], start) and np.allclose(P[-1], goal)
        assert orc.feasible_batch(P).all()
        vis, _ = orc.edges_visible_batch(P[:-1], P[1:], eps=space.eps)
        assert vis.all(), kind
        st = plan.getStats()
        assert st["feasible_samples"] > 0 and st["edges_checked"] > 0
        plan.close()


def test_triangle_primitive_and_distance_point(built):
    """GeometricPrimitive "Triangle" (one of the common primitives of Cpp/docs/Manual-Geometry.md:241-250) and
    Geometry3D.distance_point (Python/klampt/src/geometry.h:1011-1030), against the oracle and closed forms."""
    from klampt_b200 import so3
    from klampt_b200.robotsim import Geometry3D, GeometricPrimitive, TriangleMesh
    from oracle import oracle as ko
    rng = np.random.default_rng(5)
    v, t = synth.unit_cube()
    cube = Geometry3D(); cube.setTriangleMesh(TriangleMesh(v, t))
    tri = Geometry3D(); p = GeometricPrimitive(); p.setTriangle([0, 0, 0], [1, 0, 0], [0, 1, 0]); tri.setGeometricPrimitive(p)
    assert tri.type() == "GeometricPrimitive" and tri.numElements() == 1
    # a triangle poking through the cube's top face collides; lifted clear of it, the distance is the gap
    tri.setCurrentTransform(so3.identity(), [0.25, 0.25, 1.0 - 1e-3])
    assert tri.collides(cube) and cube.collides(tri)
    tri.setCurrentTransform(so3.identity(), [0.25, 0.25, 1.5])
    assert not tri.collides(cube)
    assert abs(tri.distance(cube).d - 0.5) < 1e-12 and tri.withinDistance(cube, 0.5 + 1e-9) and not tri.withinDistance(cube, 0.5 - 1e-9)
    # random poses against the oracle's triangle-triangle routines
    a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    other = Geometry3D(); q = GeometricPrimitive(); q.setTriangle(*a); other.setGeometricPrimitive(q)
    for _ in range(40):
        R = synth._random_rotation(rng); tt = rng.uniform(-1.2, 1.2, size=3)
        tri.setCurrentTransform(so3.from_matrix(R), list(tt))
        other.setCurrentTransform(so3.identity(), [0, 0, 0])
        b = a @ R.T + tt
        want_hit = ko.tri_tri_intersect(b, a)
        assert tri.collides(other) == want_hit
        if not want_hit:
            assert abs(tri.distance(other).d - ko.tri_tri_distance(b, a)) < 1e-9
    # distance_point: outside the unit cube the distance to the surface has a closed form; batched = one by one
    cube.setCurrentTransform(so3.identity(), [0, 0, 0])
    P = rng.uniform(-1.5, 2.5, size=(200, 3))
    outside = np.linalg.norm(np.maximum(np.maximum(-P, P - 1.0), 0.0), axis=1)
    inside = np.minimum(P, 1.0 - P).min(axis=1)
    want = np.where(outside > 0, outside, np.maximum(inside, 0.0))      # distance to the surface mesh, not a solid
    got = cube.distance_points_batch(P)
    np.testing.assert_allclose(got, want, atol=1e-12)
    assert abs(cube.distance_point(list(P[0])).d - want[0]) < 1e-12


def test_distance_query_state_machine_batch(built):
    """DistanceQuery's Far / Close / Contact cycle (reference Cpp/Planning/DistanceQuery.cpp:15-76) over a batch of poses:
    a unit cube approaching another along x."""
    from klampt_b200 import so3
    from klampt_b200.distancequery import DistanceQueryBatch, FAR, CLOSE, CONTACT
    from klampt_b200.robotsim import Geometry3D, TriangleMesh
    v, t = synth.unit_cube()
    a = Geometry3D(); a.setTriangleMesh(TriangleMesh(v, t))
    b = Geometry3D(); b.setTriangleMesh(TriangleMesh(v, t))
    gaps = np.array([1.0, 0.5, 0.15, 0.05, -0.2, 0.35])
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    Ta = np.tile(I, (len(gaps), 1)); Tb = np.tile(I, (len(gaps), 1)); Tb[:, 9] = 1.0 + gaps
    q = DistanceQueryBatch(a, b, len(gaps))
    d = q.UpdateQuery(Ta, Tb)
    assert list(q.s) == [FAR, FAR, CLOSE, CLOSE, CONTACT, FAR]
    np.testing.assert_allclose(d, [0.2, 0.2, 0.15, 0.05, 0.0, 0.2], atol=1e-12)
    assert q.launches == 2                                           # one within-distance launch, one distance launch
    np.testing.assert_allclose(q.UpdateQuery(Ta, Tb), d)             # same cycle: cached
    q.NextCycle()
    Tb[:, 9] = 1.0 + gaps - 0.1                                      # everything moves 10 cm closer
    d2 = q.UpdateQuery(Ta, Tb)
    assert list(q.s) == [FAR, FAR, CLOSE, CONTACT, CONTACT, FAR]
    np.testing.assert_allclose(d2, [0.2, 0.2, 0.05, 0.0, 0.0, 0.2], atol=1e-12)


def test_robot_cspace_named_tests(setup):
    """RobotCSpace.testFeasibility / feasibilityFailures with the reference's test names (plan/robotcspace.py:31-75)"""
    spec, world, collider, space, orc = setup
    names = space.feasibilityTestNamesList()
    assert names[:4] == ["joint limits", "setconfig", "calcbb", "self collision"] and sum(n.startswith("obj collision") for n in names) == 10
    Q = synth.sample_configs(spec.robot, 300, 91)
    want = orc.feasible_batch(Q)
    for i in range(40):
        x = list(Q[i])
        fails = space.feasibilityFailures(x)
        assert (len(fails) == 0) == bool(want[i])
        for n in names:
            assert space.testFeasibility(n, x) == (n not in fails)
    with pytest.raises(ValueError):
        space.testFeasibility("no such test", list(Q[0]))


def test_one_joint_limit_contract_for_single_and_batch_queries(built):
    """The reference's Python RobotCSpace tests every dimension against `bound`; the C++ space (and the engine) only Normal / Weld joints.
    The mirror answers isFeasible, feasible and feasible_batch / visible_batch by ONE rule: a floating base with a finite box is held to it
    everywhere, and a box tightened with setBounds after construction is honoured by the batch forms too."""
    from klampt_b200.collide import WorldCollider
    from klampt_b200.robotcspace import RobotCSpace
    from klampt_b200.robotsim import WorldModel
    spec = synth.world_floating(n_obstacles=6)
    world = WorldModel.from_spec(spec)
    robot = world.robot(0)
    qmin, qmax = robot.getJointLimits()
    qmin, qmax = list(qmin), list(qmax)
    for k in range(3):                                   # a finite box for the floating base's translation (the engine does not check it)
        qmin[k], qmax[k] = -0.4, 0.4
    robot.setJointLimits(qmin, qmax)
    space = RobotCSpace(robot, WorldCollider(world))
    rng = np.random.default_rng(2)
    lo = np.array([b[0] if np.isfinite(b[0]) else -3.0 for b in space.bound]); hi = np.array([b[1] if np.isfinite(b[1]) else 3.0 for b in space.bound])
    Q = rng.uniform(lo, hi, size=(3000, len(lo)))
    Q[::3, 0] += 0.6                                      # a third of the rows leave the box along x
    got = space.feasible_batch(Q)
    inbox = ((Q >= np.array([b[0] for b in space.bound])) & (Q <= np.array([b[1] for b in space.bound]))).all(axis=1)
    assert not got[~inbox].any() and got[inbox].any() and (~inbox).sum() > 300
    for i in list(np.nonzero(~inbox)[0][:20]) + list(np.nonzero(inbox)[0][:40]):
        assert bool(got[i]) == space.isFeasible(list(Q[i])) == space.feasible(list(Q[i]))
    # an edge that starts outside the box is not visible although the engine alone would pass it
    ok = Q[got == 1]
    out = ok[:50].copy(); out[:, 0] = 0.9
    vis = space.visible_batch(out, ok[50:100], eps=0.05)
    assert not vis.any()
    assert space.engine.edges_visible_batch(out, ok[50:100], eps=0.05, return_nchecks=False).any()       # the engine by itself ignores that box
    # tightening the box after construction reaches the batch forms
    b = list(space.bound); b[1] = (-0.1, 0.1); space.setBounds(b)
    got2 = space.feasible_batch(Q)
    assert not got2[np.abs(Q[:, 1]) > 0.1].any() and np.array_equal(got2[np.abs(Q[:, 1]) <= 0.1], got[np.abs(Q[:, 1]) <= 0.1])


def test_two_robots_the_other_one_is_an_obstacle(built):
    """SingleRobotCSpace checks its robot against every other robot of the world (RobotCSpace.cpp:794-823): a second arm, standing
    0.9 m away at its own configuration, takes feasible configurations away from the first; the oracle on the same description
    agrees bit for bit, and the named tests report the other robot"""
    from klampt_b200.collide import WorldCollider
    from klampt_b200.robotcspace import RobotCSpace
    from klampt_b200.robotsim import WorldModel
    from oracle.oracle import OracleWorld
    spec = synth.world_c1()
    world = WorldModel.from_spec(spec)
    base = spec.robot.T0.copy()
    second = synth.world_c1().robot
    second.T0 = base.copy()
    second.T0[0, 9:12] += np.array([0.9, 0.25, 0.0])                 # the second arm's base stands beside the first
    other = world.addRobot("second", second, spec.geoms)
    other.setConfig(list(synth.sample_configs(second, 1, 3)[0]))      # FK on the GPU places its links
    space = RobotCSpace(world.robot(0), WorldCollider(world))
    alone = RobotCSpace(WorldModel.from_spec(spec).robot(0), WorldCollider(WorldModel.from_spec(spec)))
    Q = synth.sample_configs(spec.robot, 20000, 31)
    f2, f1 = space.feasible_batch(Q), alone.feasible_batch(Q)
    assert np.all(f2 <= f1) and (f2 < f1).sum() > 50
    assert np.array_equal(f2, OracleWorld(space.spec).feasible_batch(Q))
    lost = np.flatnonzero(f2 < f1)[:5]
    for i in lost:
        assert "robot collision 1 second" in space.feasibilityFailures(list(Q[i]))
        assert not space.testFeasibility("robot collision 1 second", list(Q[i]))
    assert "robot collision 1 second" in space.feasibilityTestNamesList()
    # the id translation: first colliding pairs named in the world's numbering point at links of the second robot
    ok, pairs = space.engine.feasible_batch(Q[lost], return_pairs=True)
    wid = space.spec.world_ids[pairs]
    lo1, hi1 = world.robotLinkID(1, 0), world.robotLinkID(1, second.L - 1)
    assert any(any(lo1 <= int(x) <= hi1 for x in row) for row in wid)
