"""Vectors produced by the reference's own pure-Python rotation code (tests/golden/make_reference_so3.py imports
/root/reference/Python/klampt/math/so3.py and se3.py; the .npz travels) against this repo's mirror of those conventions, the URDF
loader's roll-pitch-yaw, and the ORACLE's SO(3) arithmetic for Floating / BallAndSocket joints (FK of the z-y-x link triple,
geodesic interpolation, angle metric) -- the one part of the oracle that real reference code can pin."""
import math
import os

import numpy as np
import pytest

from klampt_b200 import io as kio, so3, synth
from oracle.oracle import OracleWorld

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_so3.npz"))


def M(R9):
    """Klamp't column-major 9-list -> 3x3"""
    return np.asarray(R9, dtype=np.float64).reshape(3, 3).T


def test_so3_mirror_matches_reference_outputs():
    for i in range(len(G["R"])):
        R = list(G["R"][i])
        np.testing.assert_allclose(so3.matrix(R), M(R), atol=0)
        # the reference rounds rotation vectors shorter than its length threshold to the identity
        np.testing.assert_allclose(so3.exp(G["w"][i]), M(R), atol=1e-14 if np.linalg.norm(G["w"][i]) > 1e-6 else 1e-8)
        # equal as angles: the reference returns 2 pi where this mirror returns 0 (its acos branch maps sin = 0 to 2 pi - 0);
        # acos near 1 costs the reference half its digits, hence 1e-7
        d = np.asarray(so3.rpy(R)) - G["rpy"][i]
        assert np.abs(np.arctan2(np.sin(d), np.cos(d))).max() < 1e-7
        assert abs(so3.angle(M(R)) - G["angle"][i]) < 1e-7                    # acos near 0 / pi amplifies rounding
        if 1e-6 < G["angle"][i] < math.pi - 1e-3:
            np.testing.assert_allclose(so3.log(M(R)), G["moment"][i], atol=1e-9)
        np.testing.assert_allclose(so3.from_quaternion(list(G["quat"][i])), R, atol=1e-12)
        np.testing.assert_allclose(so3.inv(R), list(M(R).reshape(-1)), atol=0)                # inverse = transpose: row-major of R
        w = G["w"][i]; th = np.linalg.norm(w)
        if th > 1e-6:
            np.testing.assert_allclose(so3.from_axis_angle((list(w / th), th)), R, atol=1e-12)
    for k in range(len(G["ia"])):
        Ra, Rb = list(G["R"][G["ia"][k]]), list(G["R"][G["ib"][k]])
        np.testing.assert_allclose(so3.mul(Ra, Rb), G["mul"][k], atol=1e-14)
        ta = G["t"][G["ia"][k]]
        np.testing.assert_allclose(np.asarray(so3.apply(Ra, list(G["p"][k]))) + ta, G["applied"][k], atol=1e-14)
        T12 = so3.to_rowmajor12(Ra, list(ta))                                               # the engine's transform layout
        np.testing.assert_allclose(synth.transform_points(T12, G["p"][k][None])[0], G["applied"][k], atol=1e-14)
        R2, t2 = so3.from_rowmajor12(T12)
        np.testing.assert_allclose(R2, Ra, atol=0); np.testing.assert_allclose(t2, ta, atol=0)


def test_urdf_rpy_matches_reference_from_rpy():
    for i in range(len(G["rpy"])):
        r, p, y = G["rpy"][i]
        np.testing.assert_allclose(kio._rpy_matrix(r, p, y), M(G["R_from_rpy"][i]), atol=1e-14)
        np.testing.assert_allclose(M(G["R_from_rpy"][i]), M(G["R"][i]), atol=1e-6)      # the reference's own round trip (its acos loses digits)


@pytest.fixture(scope="module")
def floating():
    w = synth.world_floating()
    return w, OracleWorld(w)


def _zyx_of(R9):
    """Euler ZYX triplet (a about z, b about y, c about x) of a reference rotation: rpy = (roll c, pitch b, yaw a)"""
    r, p, y = so3.rpy(list(R9))
    return np.array([y, p, r])


def test_oracle_floating_fk_is_reference_from_rpy(floating):
    """the z, y, x revolute links of a floating joint compose to the reference's from_rpy((c, b, a))"""
    w, o = floating
    q = np.zeros(w.robot.L)
    for i in range(len(G["R"])):
        q[3:6] = _zyx_of(G["R"][i])
        T = o.fk(q)
        np.testing.assert_allclose(T[5][:9].reshape(3, 3), M(G["R"][i]), atol=1e-9)


def test_oracle_geodesic_interpolation_and_metric_match_reference_so3(floating):
    """Interpolate.cpp:16-52 / :229-278 on the Euler triplets of a floating joint = so3.interpolate / so3.distance of the reference"""
    w, o = floating
    a, b = np.zeros(w.robot.L), np.zeros(w.robot.L)
    checked = 0
    for k in range(len(G["ia"])):
        Ra, Rb = G["R"][G["ia"][k]], G["R"][G["ib"][k]]
        a[3:6], b[3:6] = _zyx_of(Ra), _zyx_of(Rb)
        assert abs(o.cspace_distance(a, b) - G["dist"][k]) < 1e-7
        if G["dist"][k] > math.pi - 1e-3:
            continue                                                   # antipodal pair: the geodesic is not unique
        m = o.interpolate(a, b, float(G["u"][k]))
        np.testing.assert_allclose(so3.euler_zyx_matrix(*m[3:6]), M(G["interp"][k]), atol=1e-8)
        np.testing.assert_allclose(so3.euler_zyx_matrix(*so3.euler_zyx_interp(a[3:6], b[3:6], float(G["u"][k]))), M(G["interp"][k]), atol=1e-8)
        checked += 1
    assert checked > 150


# ------------------------------------------------------------------------------------------------ collision masks
MASKS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_masks.npz"))
_KIND = {"TerrainModel": 0, "RigidObjectModel": 1, "RobotModelLink": 2}


def _worlds():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_mask", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_mask.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.worlds()


def _rows(col):
    rows = set()
    for i, s in enumerate(col.mask):
        for j in s:
            a, b = col.geomList[i][0], col.geomList[j][0]
            rows.add((_KIND[type(a).__name__], a.index, _KIND[type(b).__name__], b.index))
    return np.array(sorted(rows), dtype=np.int32).reshape(-1, 4)


@pytest.mark.parametrize("name", ["c1", "c3", "boxes", "c1_empty"])
def test_world_collider_mirror_equals_the_reference_world_collider(name):
    """klampt_b200.collide.WorldCollider against the masks the reference's own WorldCollider.__init__ / ignoreCollision produced on the
    same worlds (tests/golden/make_reference_mask.py)"""
    from klampt_b200 import robotsim
    from klampt_b200.collide import WorldCollider
    world = robotsim.WorldModel.from_spec(_worlds()[name])
    col = WorldCollider(world)
    assert np.array_equal(_rows(col), MASKS[name])
    robot = world.robot(0)
    col.ignoreCollision(robot.link(robot.numLinks() - 1))
    if world.numRigidObjects() > 0 and world.rigidObject(0).geometry().type() != "":
        col.ignoreCollision((robot.link(1), world.rigidObject(0)))
    assert np.array_equal(_rows(col), MASKS[name + "_ignored"])


@pytest.mark.parametrize("name", ["c1", "c3", "boxes", "c1_empty"])
def test_oracle_default_mask_agrees_with_the_reference_world_collider(name):
    """the C++ side's InitializeDefault mask (PlannerSettings.cpp:16-41, restated in the oracle) and the Python side's WorldCollider
    must enable the same (link, terrain), (link, object) and (link, link) pairs -- as far as the robot's feasibility test reads them:
    CheckCollision tests collisionEnabled(i,j) || collisionEnabled(j,i) for the environment and (i<j) for self pairs, and only ever
    reaches bodies with a non-empty geometry"""
    spec = _worlds()[name]
    o = OracleWorld(spec)
    m = o.pair_mask()
    T, O, L = len(spec.terrains), len(spec.objects), spec.robot.L
    lid = lambda j: T + O + 1 + j
    nonempty = lambda gi: gi >= 0 and spec.geoms[gi].kind != "empty"
    want = set()
    for j in range(L):
        if not nonempty(spec.robot.link_geom[j]):
            continue
        for t in range(T):
            if nonempty(spec.terrains[t]) and (m[lid(j), t] or m[t, lid(j)]):
                want.add((2, j, 0, t)); want.add((0, t, 2, j))
        for k in range(O):
            if nonempty(spec.objects[k][0]) and (m[lid(j), T + k] or m[T + k, lid(j)]):
                want.add((2, j, 1, k)); want.add((1, k, 2, j))
        for k in range(j + 1, L):
            if nonempty(spec.robot.link_geom[k]) and m[lid(j), lid(k)]:
                want.add((2, j, 2, k)); want.add((2, k, 2, j))
    ref = set(map(tuple, MASKS[name].tolist()))
    ref_robot = {r for r in ref if r[0] == 2 or r[2] == 2}                # the robot's pairs (object-object / terrain-object ones are the world's)
    assert want == ref_robot


# ------------------------------------------------------------------------------------------------ RobotCSpace test list
import json

CSPACE = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cspace.json")))


def _host_only_space(world, with_collider):
    """RobotCSpace without its engine (no GPU here): the host-side members the reference's constructor also produces"""
    from klampt_b200.collide import WorldCollider
    from klampt_b200.cspace import CSpace
    from klampt_b200.robotcspace import RobotCSpace
    sp = RobotCSpace.__new__(RobotCSpace)
    CSpace.__init__(sp)
    sp.robot = world.robot(0)
    sp.collider = WorldCollider(world) if with_collider else None
    sp.setBounds(list(zip(*sp.robot.getJointLimits())))
    sp.properties["geodesic"] = 1
    sp.joint_limit_failures = [0] * len(sp.bound)
    return sp


@pytest.mark.parametrize("name", sorted(k for k in CSPACE if not k.endswith("_cspace") and not k.startswith(("inactive_", "affine_"))))
def test_robot_cspace_test_list_equals_the_reference_constructor(name):
    """names, order and dependencies of the feasibility tests, bounds, properties and eps of the reference's own
    RobotCSpace.__init__ run on the same worlds (tests/golden/make_reference_cspace.py)"""
    from klampt_b200 import robotsim
    want = CSPACE[name]
    with_collider = not name.endswith("_nocollider")
    world = robotsim.WorldModel.from_spec(_worlds()[name.replace("_nocollider", "")])
    sp = _host_only_space(world, with_collider)
    assert sp.feasibilityTestNamesList() == want["names"]
    assert sorted(map(tuple, sp.feasibilityTestDependenciesList())) == sorted(map(tuple, want["dependencies"]))
    np.testing.assert_allclose(np.array(sp.bound), np.array(want["bound"]), atol=0)
    assert sp.eps == want["eps"]
    assert set(sp.properties) == set(want["properties"])
    for k, v in want["properties"].items():
        np.testing.assert_allclose(np.asarray(sp.properties[k], dtype=np.float64), np.asarray(v, dtype=np.float64), rtol=1e-15)
    lo, hi = [b[0] for b in sp.bound], [b[1] + 1e-9 for b in sp.bound]
    assert [bool(sp.inJointLimits(lo)), bool(sp.inJointLimits(hi))] == want["in_limits"]


# ------------------------------------------------------------------------------------------------ a .rob file the reference wrote
def test_rob_loader_reads_the_file_the_reference_generator_writes():
    """tests/golden/ref_planar_3R.rob is the output of the reference's model/create/planar_robot.py (make_reference_rob.py).  It has no
    `parents` line (serial chain, Robot.cpp:899-902), and its TParent line ends in a literal backslash-n that glues the `axis` line
    onto it, so -- as in the reference's reader (Robot.cpp:271-300, 409-416) -- the axes are never read and default to z
    (Robot.cpp:953-954): the "planar" robot the reference actually loads turns in the x-y plane."""
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_planar_3R.rob")).read()
    assert "0.5 0 0\\naxis" in text and "parents" not in text
    world, r = kio.parse_rob(text)
    assert r.L == 3 and list(r.parents) == [-1, 0, 1]
    np.testing.assert_allclose(r.axis, [[0, 0, 1]] * 3)
    np.testing.assert_allclose(r.T0[:, 9:], [[0, 0, 0], [0.5, 0, 0], [0.5, 0, 0]])
    np.testing.assert_allclose(r.qmin, 0.0); np.testing.assert_allclose(r.qmax, 6.28319)
    from klampt_b200.worldspec import JOINT_SPIN
    assert (r.joint_type == JOINT_SPIN).all() and all(g >= 0 for g in r.link_geom)
    g = world.geoms[r.link_geom[1]]
    assert g.tris.shape == (12, 3) and g.verts[:, 0].max() == pytest.approx(0.5)       # geomscale 0.5 on the unit-length box
    o = OracleWorld(world)
    q = [0.3, -0.2, 0.5]
    T = o.fk(q)
    np.testing.assert_allclose(T[2][9:], [0.5 * math.cos(0.3) + 0.5 * math.cos(0.1), 0.5 * math.sin(0.3) + 0.5 * math.sin(0.1), 0.0], atol=1e-15)
    c, s = math.cos(0.6), math.sin(0.6)
    np.testing.assert_allclose(T[2][:9].reshape(3, 3), [[c, -s, 0], [s, c, 0], [0, 0, 1]], atol=1e-15)


# ------------------------------------------------------------------------------------------------ bb helpers / group iterators
GROUPS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_groupiter.json")))


@pytest.mark.parametrize("name", ["dense", "sparse", "far_groups"])
def test_group_iterators_and_bb_helpers_equal_the_reference(name):
    """the module-level helpers of model/collide.py on stand-in box geometries: same pairs in the same order as the reference's own
    functions produced (tests/golden/make_reference_groupiter.py)"""
    import importlib.util
    from klampt_b200 import collide as kc
    spec = importlib.util.spec_from_file_location("make_reference_groupiter", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_groupiter.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = GROUPS[name]
    b, g1, g2 = gen.case_inputs(name)
    G, G1, G2 = [gen.BoxGeom(*x, loose=0.01) for x in b], [gen.BoxGeom(*x, loose=0.01) for x in g1], [gen.BoxGeom(*x, loose=0.01) for x in g2]
    even = lambda i, j: (i + j) % 2 == 0
    L = lambda it: [list(p) for p in it]
    assert L(kc.self_collision_iter(G)) == want["self_all"]
    assert L(kc.self_collision_iter(G, even)) == want["self_even"]
    assert L(kc.self_collision_iter(G, [(0, 1), (2, 5), (3, 4)])) == want["self_list"]
    assert L(kc.group_collision_iter(G1, G2)) == want["group_all"]
    assert L(kc.group_collision_iter(G1, G2, even)) == want["group_even"]
    assert L(kc.group_collision_iter(G1, G2, [(0, 0), (1, 2)])) == want["group_list"]
    assert L(kc.group_subset_collision_iter(G, *want["subset"])) == want["subset_all"]
    assert L(kc.group_subset_collision_iter(G, [0, 1, 2], [5, 6, 7, 8])) == [[i, j] for i in (0, 1, 2) for j in (5, 6, 7, 8) if G[i].collides(G[j])]
    assert list(map(list, kc.bb_union(*[g.getBB() for g in G]))) == want["bb_union"]
    inter = kc.bb_intersection(G[0].getBB(), G[1].getBB())
    assert list(map(list, inter)) == want["bb_intersection"] and kc.bb_empty(inter) == want["bb_empty"]
    assert list(map(list, kc.bb_create(*[g.lo for g in G]))) == want["bb_create"]
    assert [kc.bb_contains(G[0].getBB(), g.lo) for g in G] == want["bb_contains"]
    assert list(map(list, kc.bb_create())) == GROUPS["bb_create_empty"] and kc.bb_empty(kc.bb_create())
    assert list(map(list, kc.bb_union(G[0].getBB()))) == list(map(list, G[0].getBB()))      # one box: used to raise (min of a scalar)


def test_plain_cspace_members_equal_the_reference_class():
    """the pure-Python members of the reference's CSpace base class (plan/cspace.py:76-214) on the same calls"""
    from klampt_b200.cspace import CSpace
    want = CSPACE["plain_cspace"]
    sp = CSpace()
    sp.setBounds([(0.0, 2.0), (1.0, 1.0), (-1.0, 3.0)])
    sp.addFeasibilityTest(lambda x: x[0] < 1.5)
    sp.addFeasibilityTest(lambda x: x[2] > 0.0, "positive z", dependencies=["test_0"])
    sp.addFeasibilityTest(lambda x: True, dependencies="positive z")
    assert [list(b) for b in sp.bound] == want["bound"] and sp.properties == want["properties"] and sp.eps == want["eps"]
    assert sp.feasibilityTestNames == want["names"] and [list(d) for d in sp.feasibilityTestDependencies] == want["dependencies"]
    assert [sp.inBounds(p) for p in want["probes"]] == want["inBounds"]
    assert [sp.feasible(p) for p in want["probes"]] == want["feasible"]          # named tests replace the bounds check
    assert sp.getStats() == want["stats_before_setup"] == {}
    sp.setup()
    assert sp.adaptiveQueriesEnabled() and sp.getStats()["feasible_count"] == 0   # setup() enables adaptive queries for named tests


# ------------------------------------------------------------------------------------------------ forward kinematics
FK = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fk.npz"))


def _fk_robots():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_fk", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_fk.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.robots()


@pytest.mark.parametrize("name", ["arm6", "dualarm15", "floating", "planar5R"])
def test_oracle_fk_equals_the_reference_kinematics_builder(name):
    """every link's world transform from the reference's own FK code (math/autodiff/kinematics_ad.py:407-457, run by
    tests/golden/make_reference_fk.py) against the oracle's recurrence: revolute and prismatic links, branching trees, a floating base"""
    spec = _fk_robots()[name]
    o = OracleWorld(spec)
    Q, want = FK[name + "_Q"], FK[name + "_T"]
    got = o.fk_batch(Q)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-15 * max(1.0, np.abs(want).max()) * spec.robot.L)
    assert np.abs(want[1:] - want[0]).max() > 0.1                    # the configurations do move the links


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["arm6", "dualarm15", "floating", "planar5R"])
def test_fk_kernel_equals_the_reference_kinematics_builder(name, built):
    """the FK kernel (kb_fk_batch) against the same reference-computed transforms"""
    from klampt_b200.engine import Engine
    spec = _fk_robots()[name]
    eng = Engine(spec)
    Q, want = FK[name + "_Q"], FK[name + "_T"]
    got = eng.fk_batch(Q)
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-15 * max(1.0, np.abs(want).max()) * spec.robot.L)


# ------------------------------------------------------------------------------------------------ WorldCollider iterators
ITER = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_iterators.json")))


@pytest.mark.parametrize("case", sorted(ITER))
def test_world_collider_iterators_visit_the_reference_pairs(case):
    """which pairs collisionTests / collisions / robotSelfCollisions / robotObjectCollisions / robotTerrainCollisions /
    objectTerrainCollisions / objectObjectCollisions visit and report, their order and orientation: the reference's own methods on
    the same worlds with the same stand-in boxes for getBB / collides (tests/golden/make_reference_iterators.py)"""
    import importlib.util
    from klampt_b200 import robotsim
    from klampt_b200.collide import WorldCollider
    spec = importlib.util.spec_from_file_location("make_reference_iterators", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_iterators.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    name, seed = case.split("/")
    world = robotsim.WorldModel.from_spec(_worlds()[name])
    gen.stub_geometries(world, int(seed))
    got = gen.run_iterators(WorldCollider(world), world)
    want = ITER[case]
    assert sorted(got) == sorted(want)
    selfpairs = 0
    for k in want:
        if k.startswith(("collisionTests", "collisions")) and isinstance(want[k], list):
            # order inside a mask row is set order, not part of the contract.  Two documented differences: the reference pairs rigid
            # objects with themselves (mask quirk), and with two filters it lists every pair from both sides
            ref = [p for p in want[k] if p[0] != p[1]]
            selfpairs += len(want[k]) - len(ref)
            assert all(p[0][0] == "RigidObjectModel" for p in want[k] if p[0] == p[1])
            if "_vs_" in k:
                assert len(ref) == 2 * len({str(p) for p in ref})
                assert {str(p) for p in got[k]} == {str(p) for p in ref} and len(got[k]) == len({str(p) for p in got[k]}), k
            else:
                assert sorted(map(str, got[k])) == sorted(map(str, ref)), k
        elif k == "collisionTests_nobb":
            pass
        elif k == "objectObjectCollisions":
            # the (i, i) calls: the reference reports an object against itself where its mask has the (o, o) entry
            assert [x for x in got[k] if x[0] != x[1]] == [x for x in want[k] if x[0] != x[1]], k
        else:
            assert got[k] == want[k], k
    nobj = sum(1 for _, o in [(0, 0)] for _ in range(world.numRigidObjects()))
    if world.numTerrains() > 0 and nobj > 0:
        assert selfpairs > 0          # the quirk is really there in the reference's output


def test_embedded_cspace_equals_the_reference_class():
    """klampt_b200.cspaceutils.EmbeddedCSpace against the reference's own class on the same ambient space (plan/cspaceutils.py:108-203),
    plus the batch forms the batched planners use"""
    from klampt_b200.cspace import CSpace
    from klampt_b200.cspaceutils import EmbeddedCSpace
    want = CSPACE["embedded_cspace"]
    base = CSpace()
    base.setBounds([(0.0, 2.0), (1.0, 1.0), (-1.0, 3.0)])
    base.addFeasibilityTest(lambda x: x[0] < 1.5)
    base.addFeasibilityTest(lambda x: x[2] > 0.0, "positive z", dependencies=["test_0"])
    base.addFeasibilityTest(lambda x: True, dependencies="positive z")
    base.distance = lambda a, b: sum(abs(p - q) for p, q in zip(a, b))
    base.interpolate = lambda a, b, u: [p + u * (q - p) for p, q in zip(a, b)]
    emb = EmbeddedCSpace(base, [2, 0], xinit=[0.25, 1.0, 0.5])
    P = want["probes"]
    assert [list(b) for b in emb.bound] == want["bound"] and emb.eps == want["eps"]
    assert emb.feasibilityTestNames == want["names"] and [list(d) for d in emb.feasibilityTestDependencies] == want["dependencies"]
    assert [emb.lift(p) for p in P] == want["lift"] and emb.project([9.0, 8.0, 7.0]) == want["project"]
    assert [emb.feasible(p) for p in P] == want["feasible"]
    assert [[bool(f(p)) for f in emb.feasibilityTests] for p in P] == want["tests"]
    assert emb.distance(P[0], P[2]) == want["distance"] and emb.interpolate(P[0], P[2], 0.25) == want["interpolate"]
    assert emb.liftPath(P[:2]) == want["liftPath"] and emb.projectPath([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]) == want["projectPath"]
    assert EmbeddedCSpace(base, [1]).lift([0.75]) == want["default_xinit_lift"]
    with pytest.raises(ValueError, match="Invalid length of embedded space vector"):
        emb.lift([1.0])
    with pytest.raises(ValueError, match="Invalid length of ambient space vector"):
        emb.project([1.0])
    np.testing.assert_array_equal(emb.lift_batch(P), np.array(want["lift"]))
    np.testing.assert_array_equal(emb.project_batch(want["lift"]), np.array(P))


def test_batched_planner_on_an_embedded_space():
    """MotionPlan over EmbeddedCSpace: a 3-D ambient space whose middle DOF is fixed; the batch calls reach the ambient space lifted"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_plan_cpu import DiskSpace
    from klampt_b200.cspaceutils import EmbeddedCSpace
    from klampt_b200.plan import MotionPlan

    class Slab(DiskSpace):                      # the disk world in dims (0, 2); dim 1 must stay at 0.3
        def __init__(self):
            DiskSpace.__init__(self)
            self.bound = [(0.0, 1.0), (0.0, 1.0), (0.0, 1.0)]
            self.feasibilityTests, self.properties = None, {}
            self.seen = []
        def feasible_batch(self, Q):
            Q = np.atleast_2d(Q)
            self.seen.append(Q[:, 1].copy())
            return (DiskSpace.feasible_batch(self, Q[:, [0, 2]]).astype(bool) & (np.abs(Q[:, 1] - 0.3) < 1e-12)).astype(np.uint8)
        def visible_batch(self, A, B):
            out = np.ones(len(A), dtype=np.uint8)
            for u in np.linspace(0, 1, 101)[1:-1]:
                out &= self.feasible_batch(A * (1 - u) + B * u)
            return out

    amb = Slab()
    emb = EmbeddedCSpace(amb, [0, 2], xinit=[0.0, 0.3, 0.0])
    MotionPlan.setOptions(knn=8, batch=200, seed=3)
    plan = MotionPlan(emb, "prm")
    plan.setEndpoints([0.05, 0.5], [0.95, 0.5])
    path = None
    for _ in range(10):
        plan.planMore(1)
        path = plan.getPath()
        if path:
            break
    assert path is not None and len(path[0]) == 2
    amb_path = np.array(emb.liftPath(path))
    assert (amb_path[:, 1] == 0.3).all() and amb.visible_batch(amb_path[:-1], amb_path[1:]).all()
    assert all((np.abs(s - 0.3) < 1e-12).all() for s in amb.seen)          # the fixed DOF never left its value


@pytest.mark.parametrize("key", sorted(k for k in CSPACE if k.startswith("inactive_")))
def test_disable_inactive_collisions_equals_the_reference(key):
    """robotplanning.disable_inactive_collisions against EmbeddedRobotCSpace.disableInactiveCollisions of the reference
    (plan/robotcspace.py:365-392) on the arm, for four moving subsets -- including its root-link lookup quirk"""
    from klampt_b200 import robotsim
    from klampt_b200.collide import WorldCollider
    from klampt_b200.robotplanning import disable_inactive_collisions
    subset = [int(x) for x in key.split("_")[1:]]
    world = robotsim.WorldModel.from_spec(_worlds()["c1"])
    col = WorldCollider(world)
    before = len(_rows(col))
    disable_inactive_collisions(col, world.robot(0), subset)
    assert _rows(col).tolist() == CSPACE[key]["mask"]
    # moving the base link moves everything; and through the reference's active[-1] lookup so does moving the LAST link
    assert (len(_rows(col)) < before) == (0 not in subset and 6 not in subset)


class _FakeEngine:
    """stands in for the CUDA engine in host-logic tests: feasible iff q[0] < 1; an edge is visible iff both ends are feasible"""
    def feasible_batch(self, Q, return_pairs=False):
        Q = np.asarray(Q, dtype=np.float64).reshape(-1, 2)
        ok = (Q[:, 0] < 1.0).astype(np.uint8)
        return (ok, np.full((len(ok), 2), -1, dtype=np.int32)) if return_pairs else ok

    def edges_visible_batch(self, A, B, eps=0.01, return_nchecks=False):
        A, B = np.asarray(A, dtype=np.float64).reshape(-1, 2), np.asarray(B, dtype=np.float64).reshape(-1, 2)
        vis = ((A[:, 0] < 1.0) & (B[:, 0] < 1.0)).astype(np.uint8)
        return (vis, np.zeros(len(vis), dtype=np.int32)) if return_nchecks else vis


def test_user_constraints_reach_the_batch_calls():
    """RobotCSpace.addConstraint: the extra predicate is applied to the rows the engine passes, and walked along the edges the engine
    found visible, at the space's resolution"""
    from klampt_b200.cspace import CSpace
    from klampt_b200.robotcspace import RobotCSpace
    sp = RobotCSpace.__new__(RobotCSpace)
    CSpace.__init__(sp)
    sp.setBounds([(0.0, 2.0), (0.0, 2.0)])
    sp.engine, sp._extra, sp.eps = _FakeEngine(), [], 0.05
    sp.distance = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)))
    sp.interpolate = lambda a, b, u: [p + u * (q - p) for p, q in zip(a, b)]
    Q = np.array([[0.5, 0.5], [0.5, 1.5], [1.5, 0.5]])
    assert list(sp.feasible_batch(Q)) == [1, 1, 0]
    calls = []
    sp.addConstraint(lambda q: calls.append(tuple(q)) or not (0.9 < q[1] < 1.1), "no band")
    assert list(sp.feasible_batch(Q)) == [1, 1, 0] and len(calls) == 2            # the infeasible row never reaches the callable
    assert list(sp.feasible_batch(np.array([[0.5, 1.0]]))) == [0] and not sp.feasible([0.5, 1.0]) and sp.feasible([0.5, 0.2])
    A = np.array([[0.2, 0.2], [0.2, 0.2], [0.2, 0.2]])
    B = np.array([[0.8, 0.8], [0.8, 1.8], [1.8, 0.2]])
    vis, n = sp.visible_batch(A, B, return_nchecks=True)
    assert list(vis) == [1, 0, 0]                                                  # the second edge crosses the band, the third leaves the engine's set
    assert sp.visible([0.2, 1.2], [0.8, 1.8]) and not sp.visible([0.2, 0.2], [0.2, 1.8])
    assert "no band" in sp.feasibilityTestNames


def test_plan_to_config_host_logic(monkeypatch):
    """robotplanning.plan_to_config: 'auto' moving subset, the fixed-DOF error, the embedded plan speaking ambient configurations"""
    from klampt_b200 import robotplanning, robotsim
    from klampt_b200.cspace import CSpace
    import klampt_b200.robotcspace as rcs

    class HostSpace(CSpace):                       # RobotCSpace without the engine: everything is feasible inside the limits
        def __init__(self, robot, collider=None, device=0):
            CSpace.__init__(self)
            self.robot, self.collider = robot, collider
            self.setBounds(list(zip(*robot.getJointLimits())))
        def addConstraint(self, c, name=None):
            self.addFeasibilityTest(c, name)
        def feasible_batch(self, Q):
            Q = np.asarray(Q, dtype=np.float64).reshape(-1, len(self.bound))
            lo, hi = np.array(self.bound).T
            return ((Q >= lo) & (Q <= hi)).all(axis=1).astype(np.uint8)
        def visible_batch(self, A, B):
            return self.feasible_batch(A) & self.feasible_batch(B)
    monkeypatch.setattr(rcs, "RobotCSpace", HostSpace)
    world = robotsim.WorldModel.from_spec(_worlds()["c1"])
    robot = world.robot(0)
    q0 = np.zeros(robot.numLinks()); robot._q = q0.copy()
    target = list(q0); target[2] = 0.7; target[4] = -0.4
    plan = robotplanning.plan_to_config(world, robot, target, type="prm", batch=64, knn=6, seed=1)
    assert plan.space.mapping == [2, 4] and plan.space.eps == 1e-2 and plan.space.xinit == list(q0)
    path = None
    for _ in range(5):
        plan.planMore(1)
        path = plan.getPath()
        if path:
            break
    assert path is not None and path[0] == list(q0) and path[-1] == target and all(len(q) == robot.numLinks() for q in path)
    assert all(q[k] == 0.0 for q in path for k in range(robot.numLinks()) if k not in (2, 4))
    with pytest.raises(ValueError, match="fixed DOF"):
        robotplanning.plan_to_config(world, robot, target, movingSubset=[2])
    full = robotplanning.plan_to_config(world, robot, target, movingSubset="all", type="prm")
    assert not hasattr(full.space, "lift") and len(full.space.bound) == robot.numLinks()
    bad = list(q0); bad[1] = 100.0
    with pytest.warns(UserWarning, match="Goal configuration fails"):
        assert robotplanning.plan_to_config(world, robot, bad, type="prm") is None
    with pytest.raises(NotImplementedError):
        robotplanning.make_space(world, robot, equalityConstraints=[object()])


def test_oracle_spin_joint_arithmetic_equals_reference_so2():
    """Spin joints (and the angle of FloatingPlanar joints): Interpolate.cpp's AngleInterp / AngleDiff against the reference's own
    math/so2.py interp / diff -- the oracle, and the robot-model mirror the planners' host side uses"""
    from klampt_b200 import robotsim
    from klampt_b200.worldspec import JOINT_SPIN, JOINT_NORMAL, WorldSpec
    G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_so3.npz"))
    w = WorldSpec()
    r = synth.make_planar_nR(w, 2)
    r.joint_type = np.array([JOINT_SPIN, JOINT_NORMAL], dtype=np.uint8)
    r.qmin[:], r.qmax[:] = -10.0, 10.0
    w.robot = r
    o = OracleWorld(w)
    robot = robotsim.WorldModel.from_spec(w).robot(0)
    wrap = lambda d: math.atan2(math.sin(d), math.cos(d))
    for a, b, u, dref, iref in zip(G2["ang_a"], G2["ang_b"], G2["ang_u"], G2["so2_diff"], G2["so2_interp"]):
        qa, qb = np.array([a, 0.25]), np.array([b, 0.25])
        assert abs(o.cspace_distance(qa, qb) - abs(dref)) < 1e-12                 # |so2.diff(a, b)|
        m = o.interpolate(qa, qb, float(u))
        if abs(abs(dref) - math.pi) > 1e-9:                                       # at exactly half a turn either way round is right
            assert abs(wrap(m[0] - iref)) < 1e-12 and m[1] == 0.25
            assert abs(wrap(robot.interpolate(list(qa), list(qb), float(u))[0] - iref)) < 1e-12
        assert abs(robot.distance(list(qa), list(qb)) - abs(dref)) < 1e-12


def test_rob_loader_reads_the_reference_moving_base_template(tmp_path):
    """tests/golden/ref_moving_base.rob is what the reference's model/create/moving_base_robot.py:12-137 writes around a geometry
    file (quoted link names, empty geometry strings, -inf / inf limits, continuation lines, a `property sensors <...>` line, one
    accMax entry too many).  Loaded here, its configuration follows moving_base_robot.set_xform (:148-161): q = (t, yaw, pitch, roll)
    places link 5 at (R, t) -- checked with rotations from the reference's so3."""
    kio.save_off(str(tmp_path / "cube.off"), *synth.unit_cube())
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_moving_base.rob")).read()
    world, r = kio.parse_rob(text, basedir=str(tmp_path))
    from klampt_b200.worldspec import JOINT_NORMAL, JOINT_SPIN, PRISMATIC, REVOLUTE
    assert r.L == 6 and list(r.parents) == [-1, 0, 1, 2, 3, 4] and r.names == ["tx", "ty", "tz", "rz", "ry", "rx"]
    assert list(r.linktype) == [PRISMATIC] * 3 + [REVOLUTE] * 3
    np.testing.assert_array_equal(r.axis, [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 1], [0, 1, 0], [1, 0, 0]])
    assert list(r.joint_type) == [JOINT_NORMAL] * 3 + [JOINT_SPIN] * 3
    assert list(r.qmin[:3]) == [-1, -1, -1] and np.isneginf(r.qmin[3:]).all() and np.isposinf(r.qmax[3:]).all()
    assert r.link_geom[:5] == [-1] * 5 and r.link_geom[5] >= 0 and len(r.drivers) == 6
    assert world.geoms[r.link_geom[5]].tris.shape == (12, 3)
    o = OracleWorld(world)
    for i in range(0, len(G["R"]), 7):
        roll, pitch, yaw = so3.rpy(list(G["R"][i]))
        t = np.clip(G["t"][i], -1, 1)
        T = o.fk(np.array([t[0], t[1], t[2], yaw, pitch, roll]))
        np.testing.assert_allclose(T[5][:9].reshape(3, 3), M(G["R"][i]), atol=1e-9)
        np.testing.assert_allclose(T[5][9:], t, atol=1e-15)


def test_rob_loader_mounts_a_robot_on_the_reference_moving_base(tmp_path):
    """tests/golden/ref_moving_base_mounted.rob: the reference generator's template for a robot FILE -- a floating cube with
    `mount 5 "ref_planar_3R.rob" <T> as "ref_planar_3R"` (Robot.cpp:648-690, RobotModel::Mount :1895-2007).  The mounted chain is the
    other reference-written file, so both come from Klamp't's own generators."""
    import shutil
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    shutil.copy(os.path.join(here, "ref_planar_3R.rob"), tmp_path / "ref_planar_3R.rob")
    world, r = kio.parse_rob(open(os.path.join(here, "ref_moving_base_mounted.rob")).read(), basedir=str(tmp_path))
    from klampt_b200.worldspec import JOINT_SPIN, JOINT_NORMAL
    assert r.L == 9 and list(r.parents) == [-1, 0, 1, 2, 3, 4, 5, 6, 7]
    assert r.names[5] == "rx" and r.names[6:] == ["ref_planar_3R:Link_0", "ref_planar_3R:Link_1", "ref_planar_3R:Link_2"]
    assert list(r.joint_type) == [JOINT_NORMAL] * 3 + [JOINT_SPIN] * 6 and list(r.joint_link) == list(range(9))
    assert r.link_geom[5] >= 0 and all(g >= 0 for g in r.link_geom[6:]) and len(r.drivers) == 6
    assert world.robot is r and len(r.qmin) == 9 and r.T0.shape == (9, 12) and r.axis.shape == (9, 3)
    np.testing.assert_allclose(r.T0[7, 9:], [0.5, 0, 0])
    o = OracleWorld(world)
    roll, pitch, yaw = so3.rpy(list(G["R"][9]))
    t = np.array([0.3, -0.2, 0.5])
    q = np.array([t[0], t[1], t[2], yaw, pitch, roll, 0.4, -0.3, 0.2])
    T = o.fk(q)
    Rb = M(G["R"][9])
    np.testing.assert_allclose(T[5][:9].reshape(3, 3), Rb, atol=1e-9)
    # the planar chain (axes z: see the planar test) rides on the base: link 8's origin in the base frame
    p8 = np.array([0.5 * math.cos(0.4) + 0.5 * math.cos(0.1), 0.5 * math.sin(0.4) + 0.5 * math.sin(0.1), 0.0])
    np.testing.assert_allclose(T[8][9:], Rb @ p8 + t, atol=1e-9)
    # default self collisions: every old/new pair but the mount link with the chain's root (parent / child)
    m = o.pair_mask()
    lid = lambda j: 1 + j
    assert not m[lid(5), lid(6)] and m[lid(5), lid(7)] and m[lid(5), lid(8)] and not m[lid(6), lid(7)] and m[lid(6), lid(8)]


def test_config_and_xform_text_formats_equal_the_reference_loader(tmp_path):
    """.config / .configs / .xform text as the reference's io/loader.py writes and reads it (tests/golden/make_reference_loader.py)"""
    L = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loader.json")))
    Q = np.array(L["Q"])
    assert kio.write_config(Q[0]) == L["config_text"] and kio.write_configs(Q) == L["configs_text"]
    np.testing.assert_array_equal(kio.read_configs(L["configs_text"]), Q)
    np.testing.assert_array_equal(kio.read_configs(L["configs_text"]), np.array(L["read_back_configs"]))
    np.testing.assert_array_equal(kio.read_config(L["config_text"]), Q[0])
    T12 = so3.to_rowmajor12(L["R"], L["t"])
    assert kio.write_xform(T12) == L["xform_text"]
    np.testing.assert_array_equal(kio.read_xform(L["xform_text"]), T12)
    R2, t2 = so3.from_rowmajor12(kio.read_xform(L["xform_text"]))
    np.testing.assert_allclose(R2, L["read_back_xform"][0], atol=0); np.testing.assert_allclose(t2, L["read_back_xform"][1], atol=0)
    kio.save_configs(str(tmp_path / "batch.configs"), Q)
    np.testing.assert_array_equal(kio.load_configs(str(tmp_path / "batch.configs")), Q)
    with pytest.raises(ValueError, match="Invalid number of items"):
        kio.read_config("3\t1.0 2.0")
    with pytest.raises(ValueError, match="different lengths"):
        kio.read_configs("2 1 2\n3 1 2 3")
    assert kio.read_configs("").shape == (0, 0)


def test_affine_embedded_cspace_equals_the_reference_class():
    """driver-space planning (q = A x + b): lift / least-squares project against the reference's AffineEmbeddedCSpace on the same
    matrix; the bounds here are the exact interval (offset included, all coupled links considered), inside the reference's naive one"""
    from klampt_b200.cspace import CSpace
    from klampt_b200.cspaceutils import AffineEmbeddedCSpace
    from klampt_b200.worldspec import DriverSpec
    want = CSPACE["affine_cspace"]
    amb = CSpace()
    amb.setBounds([(-1.0, 1.0), (-2.0, 2.0), (-0.5, 0.5), (0.0, 1.0)])
    A = np.array([[1.0, 0.0], [0.0, 2.0], [0.0, -1.0], [0.0, 0.0]])
    aff = AffineEmbeddedCSpace(amb, A, [0.0, 0.0, 0.1, 0.25])
    xs = [[0.5, 0.2], [-1.0, -0.4], [0.0, 0.0]]
    np.testing.assert_allclose([aff.lift(x) for x in xs], want["lift"], atol=1e-15)
    np.testing.assert_allclose([aff.project(aff.lift(x)) for x in xs], want["project"], atol=1e-12)
    np.testing.assert_allclose(aff.project([0.3, 1.0, 0.2, 0.9]), want["project_off_manifold"], atol=1e-12)
    assert aff.eps == want["eps"]
    for (lo, hi), ref in zip(aff.bound, want["bound_as_sets"]):
        assert ref[0] <= lo < hi <= ref[1]
    np.testing.assert_allclose(aff.bound, [(-1.0, 1.0), (-0.4, 0.6)], atol=1e-15)
    corners = aff.lift_batch([[lo, hi] for lo in aff.bound[0] for hi in aff.bound[1]])
    assert all(amb.inBounds(list(c)) for c in corners)                    # the whole driver box maps inside the ambient bounds
    np.testing.assert_allclose(aff.project_batch(aff.lift_batch(xs)), xs, atol=1e-12)
    # the same space from driver records: a normal driver on link 0, an affine one over links 1 and 2
    drv = [DriverSpec([0], [1.0], [0.0], -0.8, 0.8), DriverSpec([1, 2], [2.0, -1.0], [0.0, 0.1], -0.3, 5.0)]
    aff2 = AffineEmbeddedCSpace.from_drivers(amb, drv, 4)
    np.testing.assert_array_equal(aff2.A, A)
    np.testing.assert_allclose(aff2.b, [0.0, 0.0, 0.1, 0.0])
    np.testing.assert_allclose(aff2.bound, [(-0.8, 0.8), (-0.3, 0.6)], atol=1e-15)
    with pytest.raises(ValueError, match="Invalid length of embedded space vector"):
        aff.lift([1.0])


def test_make_space_plans_in_driver_space_for_coupled_links(monkeypatch):
    """robotplanning.make_space on a robot with an affine driver: the returned space is the driver-space embedding and a batched plan
    over it keeps the coupled links coupled"""
    from types import SimpleNamespace
    from klampt_b200 import robotplanning, robotsim
    from klampt_b200.cspace import CSpace
    from klampt_b200.cspaceutils import AffineEmbeddedCSpace, EmbeddedMotionPlan
    from klampt_b200.worldspec import DriverSpec
    import klampt_b200.robotcspace as rcs
    L = 7
    drivers = [DriverSpec([k], [1.0], [0.0], -2.0, 2.0) for k in range(5)] + [DriverSpec([5, 6], [1.0, -1.0], [0.0, 0.0], -1.0, 1.0)]

    class HostSpace(CSpace):
        def __init__(self, robot, collider=None, device=0):
            CSpace.__init__(self)
            self.robot, self.collider = robot, collider
            self.setBounds(list(zip(*robot.getJointLimits())))
            self.spec = SimpleNamespace(robot=SimpleNamespace(drivers=drivers))
        def addConstraint(self, c, name=None):
            self.addFeasibilityTest(c, name)
        def feasible_batch(self, Q):
            Q = np.asarray(Q, dtype=np.float64).reshape(-1, L)
            lo, hi = np.array(self.bound).T
            return (((Q >= lo) & (Q <= hi)).all(axis=1) & (np.abs(Q[:, 5] + Q[:, 6]) < 1e-12)).astype(np.uint8)     # only coupled fingers are feasible
        def visible_batch(self, A, B):
            return self.feasible_batch(A) & self.feasible_batch(B)
    monkeypatch.setattr(rcs, "RobotCSpace", HostSpace)
    world = robotsim.WorldModel.from_spec(_worlds()["c1"])
    robot = world.robot(0)
    robot._q = np.zeros(L)
    space = robotplanning.make_space(world, robot)
    assert isinstance(space, AffineEmbeddedCSpace) and space.m == 6 and space.n == 7 and space.bound[5] == (-1.0, 1.0)
    plan = EmbeddedMotionPlan(space, None, "prm", batch=128, knn=6, seed=2)
    goal = [0.0, -0.3, 0.2, 0.1, 0.0, 0.4, -0.4]                     # link 0 is the fixed base
    plan.setEndpoints([0.0] * L, goal)
    path = None
    for _ in range(6):
        plan.planMore(1)
        path = plan.getPath()
        if path:
            break
    assert path is not None and np.allclose(path[-1], goal) and all(abs(q[5] + q[6]) < 1e-12 for q in path)


def test_affine_space_from_drivers_equals_reference_from_robot_drivers():
    """AffineEmbeddedCSpace.from_drivers against the reference's fromRobotDrivers run on a mirror robot with an affine driver
    (plan/cspaceutils.py:320-369) -- and the mirror's driver accessors, which that reference code called"""
    from klampt_b200 import robotsim
    from klampt_b200.cspace import CSpace
    from klampt_b200.cspaceutils import AffineEmbeddedCSpace
    from klampt_b200.worldspec import DriverSpec
    want = CSPACE["affine_from_drivers"]
    spec = _worlds()["c1"]
    spec.robot.drivers = [DriverSpec([k], [1.0], [0.0], -2.0, 2.0) for k in range(1, 5)] + [DriverSpec([5, 6], [1.0, -0.5], [0.0, 0.2], -1.0, 1.0)]
    robot = robotsim.WorldModel.from_spec(spec).robot(0)
    assert robot.numDrivers() == 5 and robot.driver(4).getType() == "affine" and robot.driver(0).getType() == "normal"
    assert robot.driver(4).getAffectedLinks() == [5, 6] and robot.driver(4).getAffineCoeffs() == ([1.0, -0.5], [0.0, 0.2])
    amb = CSpace()
    amb.setBounds(list(zip(*robot.getJointLimits())))
    aff = AffineEmbeddedCSpace.from_drivers(amb, spec.robot.drivers, robot.numLinks())
    np.testing.assert_array_equal(aff.A, np.array(want["A"]))
    np.testing.assert_array_equal(aff.b, want["b"])
    np.testing.assert_allclose(aff.lift([0.1, 0.2, 0.3, 0.4, 0.5]), want["lift"], atol=1e-16)
