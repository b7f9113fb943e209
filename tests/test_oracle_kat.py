"""Pins the CPU oracle: analytic known-answer tests, an independent second method (numpy / scipy brute force), and
the only fixtures the reference itself offers for this path: the unit cube of tests/objects/cube.off,
the so3 known-answer of tests/test_math_so3.py:14-19 and the planar nR arm of model/create/planar_robot.py."""
import math

import numpy as np
import pytest

from klampt_b200 import so3, synth
from klampt_b200.worldspec import GeomSpec, WorldSpec, DriverSpec, JOINT_SPIN, JOINT_NORMAL
from oracle import oracle as ko
from oracle.oracle import OracleWorld

I12 = synth.IDENTITY12


# ------------------------------------------------------------------------------------------ so3 (reference KAT)
def test_so3_rpy_reference_kat():
    R = so3.from_quaternion((-4.32978e-17, -0.707107, 4.32978e-17, 0.707107))
    r, p, y = so3.rpy(R)
    assert r == pytest.approx(0.0, abs=5e-8)
    assert p == pytest.approx(1.5707963267948966, abs=5e-8)
    assert y == pytest.approx(3.141592653589793, abs=5e-8)


def test_so3_rpy_branch_representative():
    """the stated deviation of so3.rpy: angles are folded into [0, 2 pi) with atan2, so the identity answers (0, 0, 0) where the
    reference's acos / sign branches (math/so3.py:95-106) answer (2 pi, 0, 2 pi); the rotation is the same either way"""
    assert so3.rpy(so3.identity()) == (0.0, 0.0, 0.0)
    for trip in ((2 * math.pi, 0.0, 2 * math.pi), (0.3, -0.4, 6.0), (5.0, 1.2, 0.0)):
        R = so3.from_matrix(so3.euler_zyx_matrix(trip[2], trip[1], trip[0]))        # Rz(yaw) Ry(pitch) Rx(roll)
        r, p, y = so3.rpy(R)
        assert np.allclose(so3.euler_zyx_matrix(y, p, r), so3.matrix(R), atol=1e-12)
        assert 0.0 <= r < 2 * math.pi and 0.0 <= y < 2 * math.pi


def test_so3_column_major_convention():
    R = so3.from_axis_angle(((0, 0, 1), math.pi / 2))
    # column-major: first three entries are the first COLUMN = image of the x axis = +y
    assert np.allclose(R[:3], [0, 1, 0]) and np.allclose(so3.apply(R, [1, 0, 0]), [0, 1, 0])
    T = so3.to_rowmajor12(R, [1, 2, 3])
    assert np.allclose(T[:9].reshape(3, 3) @ [1, 0, 0], [0, 1, 0]) and np.allclose(T[9:], [1, 2, 3])
    R2, t2 = so3.from_rowmajor12(T)
    assert np.allclose(R2, R) and np.allclose(t2, [1, 2, 3])


# ------------------------------------------------------------------------------------------ element predicates
def tri(*p):
    return np.array(p, dtype=np.float64).reshape(9)


def test_tri_tri_known_answers():
    A = tri([0, 0, 0], [1, 0, 0], [0, 1, 0])
    assert ko.tri_tri_intersect(A, tri([0.2, 0.2, -1], [0.2, 0.2, 1], [0.8, 0.8, 1]))            # pierces
    assert not ko.tri_tri_intersect(A, tri([0.2, 0.2, 0.1], [0.3, 0.2, 1], [0.8, 0.8, 1]))        # above
    assert not ko.tri_tri_intersect(A, tri([2, 2, -1], [2, 2, 1], [3, 3, 1]))                     # crosses the plane far away
    assert ko.tri_tri_intersect(A, tri([0.25, 0.25, 0], [0.3, 0.25, 0], [0.25, 0.3, 0]))          # coplanar, contained
    assert ko.tri_tri_intersect(A, tri([0.5, 0.5, 0], [2, 2, 0], [2, 0.5, 0]))                    # coplanar, touching the hypotenuse
    assert not ko.tri_tri_intersect(A, tri([0.6, 0.6, 0], [2, 2, 0], [2, 0.6, 0]))                # coplanar, disjoint
    assert ko.tri_tri_intersect(A, tri([1, 0, 0], [2, 0, 1], [2, 0, -1]))                         # shares a vertex
    assert ko.tri_tri_intersect(A, tri([0.5, 0, 0], [0.5, -1, 1], [0.5, -1, -1]))                 # vertex on an edge
    assert ko.tri_tri_distance(A, tri([0, 0, 2], [1, 0, 2], [0, 1, 2])) == pytest.approx(2.0)
    assert ko.tri_tri_distance(A, tri([2, 0, 0], [3, 0, 0], [2, 1, 0])) == pytest.approx(1.0)     # vertex-vertex in plane
    assert ko.tri_tri_distance(A, tri([0.2, 0.2, -1], [0.2, 0.2, 1], [0.8, 0.8, 1])) == 0.0


def test_point_tri_and_seg_seg_known_answers():
    T = tri([0, 0, 0], [2, 0, 0], [0, 2, 0])
    assert ko.point_tri_distance([0.5, 0.5, 3], T) == pytest.approx(3.0)                # face region
    assert ko.point_tri_distance([-1, -1, 0], T) == pytest.approx(math.sqrt(2))         # vertex region
    assert ko.point_tri_distance([1, -2, 0], T) == pytest.approx(2.0)                   # edge region
    assert ko.point_tri_distance([2, 2, 0], T) == pytest.approx(math.sqrt(2))           # hypotenuse
    assert ko.seg_seg_distance([0, 0, 0], [1, 0, 0], [0, 1, 1], [1, 1, 1]) == pytest.approx(math.sqrt(2))
    assert ko.seg_seg_distance([0, 0, 0], [1, 0, 0], [0.5, -1, 0.5], [0.5, 1, 0.5]) == pytest.approx(0.5)   # crossing, skew
    assert ko.seg_seg_distance([0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]) == pytest.approx(1.0)           # collinear
    assert ko.seg_seg_distance([0, 0, 0], [0, 0, 0], [1, 1, 0], [1, 1, 0]) == pytest.approx(math.sqrt(2))   # degenerate


def _np_seg_tri(p, q, a, b, c):
    """independent restatement: Moller-Trumbore on the closed segment"""
    d, e1, e2 = q - p, b - a, c - a
    h = np.cross(d, e2)
    det = e1 @ h
    if abs(det) < 1e-14:
        return None
    s = p - a
    u = (s @ h) / det
    qv = np.cross(s, e1)
    v = (d @ qv) / det
    t = (e2 @ qv) / det
    return u >= 0 and v >= 0 and u + v <= 1 and 0 <= t <= 1


def test_tri_tri_against_numpy_second_method():
    rng = np.random.default_rng(11)
    n_hit = 0
    for _ in range(4000):
        A = rng.uniform(-1, 1, size=(3, 3))
        B = rng.uniform(-1, 1, size=(3, 3)) * rng.uniform(0.2, 1.5)
        want = False
        for X, Y in ((A, B), (B, A)):
            for i in range(3):
                r = _np_seg_tri(X[i], X[(i + 1) % 3], Y[0], Y[1], Y[2])
                want = want or bool(r)
        got = ko.tri_tri_intersect(A.reshape(9), B.reshape(9))
        assert got == want
        n_hit += got
    assert 400 < n_hit < 3600


def test_tri_tri_distance_against_sampling():
    rng = np.random.default_rng(12)
    w = rng.dirichlet(np.ones(3), size=4000)
    for _ in range(60):
        A = rng.uniform(-1, 1, size=(3, 3))
        B = rng.uniform(-1, 1, size=(3, 3)) + rng.uniform(-2, 2, size=3)
        d = ko.tri_tri_distance(A.reshape(9), B.reshape(9))
        pa, pb = w @ A, w @ B
        # every sampled point of A is at least d away from triangle B and vice versa, and some sample gets close
        dm = min(min(ko.point_tri_distance(p, B.reshape(9)) for p in pa[:300]), min(ko.point_tri_distance(p, A.reshape(9)) for p in pb[:300]))
        assert dm >= d - 1e-12
        assert dm <= d + 0.25


# ------------------------------------------------------------------------------------------ cube - cube analytic
@pytest.fixture(scope="module")
def cubes():
    v, t = synth.unit_cube()
    w = WorldSpec()
    ga = w.add_geom(GeomSpec.mesh(v, t))
    gb = w.add_geom(GeomSpec.mesh(v, t, margin=0.0))
    gm = w.add_geom(GeomSpec.mesh(v, t, margin=0.1))
    gs = w.add_geom(GeomSpec.sphere([0, 0, 0], 0.25))
    w.robot = synth.make_planar_nR(w, 1)
    return OracleWorld(w), ga, gb, gm, gs


def T_at(x, y=0.0, z=0.0, R=None):
    return synth.make_T(R, [x, y, z])


def test_cube_cube_analytic(cubes):
    o, ga, gb, gm, gs = cubes
    assert o.geom_distance(ga, I12, gb, T_at(1.5)) == pytest.approx(0.5, abs=1e-15)
    assert o.geom_distance(ga, I12, gb, T_at(2.0, 2.0)) == pytest.approx(math.sqrt(2), abs=1e-15)     # edge - edge
    assert o.geom_distance(ga, I12, gb, T_at(2.0, 2.0, 2.0)) == pytest.approx(math.sqrt(3), abs=1e-15)  # corner - corner
    assert not o.geom_collides(ga, I12, gb, T_at(1.5))
    assert o.geom_collides(ga, I12, gb, T_at(0.5, 0.25, 0.25))
    assert o.geom_distance(ga, I12, gb, T_at(0.5, 0.25, 0.25)) == 0.0
    # surfaces only: a small cube strictly inside the big one does NOT collide (src/geometry.h:1082-1085)
    S = np.concatenate([(0.2 * np.eye(3)).reshape(-1), [0.4, 0.4, 0.4]])
    assert not o.geom_collides(ga, I12, gb, S)
    assert o.geom_distance(ga, I12, gb, S) == pytest.approx(0.4, abs=1e-15)
    # rotated 45 deg about z, corner pointing at the face x = 1
    R = synth.rot_axis_angle([0, 0, 1], math.pi / 4)
    gap = o.geom_distance(ga, I12, gb, T_at(1.0 + math.sqrt(0.5) + 0.3, -0.2, 0.0, R))   # B rotates about its origin corner: x spans [c - sqrt(.5), c + sqrt(.5)], nearest edge at y = 0.507
    assert gap == pytest.approx(0.3, abs=1e-12)


def test_margin_semantics(cubes):
    o, ga, gb, gm, gs = cubes
    # A,B collide iff dist <= margin_A + margin_B; reported distance = geometric distance - margins (Manual-Geometry.md:17)
    assert o.geom_distance(ga, I12, gm, T_at(1.5)) == pytest.approx(0.4, abs=1e-15)
    assert not o.geom_collides(ga, I12, gm, T_at(1.11))
    assert o.geom_collides(ga, I12, gm, T_at(1.09))
    assert o.geom_within_distance(ga, I12, gb, T_at(1.5), 0.5)        # closed: dist == tol counts
    assert not o.geom_within_distance(ga, I12, gb, T_at(1.5), 0.499)
    assert o.geom_distance(ga, I12, gb, T_at(3.0), upper_bound=0.5) == 0.5   # capped at the bound


def test_sphere_primitive_signed_distance(cubes):
    o, ga, gb, gm, gs = cubes
    assert o.geom_distance(gs, T_at(2.0, 0.5, 0.5), ga, I12) == pytest.approx(0.75, abs=1e-15)
    assert o.geom_distance(gs, T_at(1.1, 0.5, 0.5), ga, I12) == pytest.approx(-0.15, abs=1e-15)   # surface cuts the ball
    assert o.geom_collides(gs, T_at(1.1, 0.5, 0.5), ga, I12)
    assert not o.geom_collides(gs, T_at(0.5, 0.5, 0.5), ga, I12)     # ball strictly inside the hollow cube
    assert o.geom_distance(gs, T_at(0, 0, 0), gs, T_at(1, 0, 0)) == pytest.approx(0.5, abs=1e-15)


def test_bvh_distance_equals_brute_force():
    rng = np.random.default_rng(5)
    w = WorldSpec()
    ga = w.add_geom(GeomSpec.mesh(*synth.blob_mesh(rng, 2, 0.3)))
    gb = w.add_geom(GeomSpec.mesh(*synth.capsule_mesh(0.1, 0, 0.4, nseg=10, ncap=3, nbody=3)))
    gc = w.add_geom(GeomSpec.cloud(rng.normal(size=(500, 3)) * 0.2, rng.uniform(0, 0.02, size=500)))
    w.robot = synth.make_planar_nR(w, 1)
    o = OracleWorld(w)
    for _ in range(25):
        Ta = synth.make_T(synth._random_rotation(rng), rng.uniform(-0.5, 0.5, size=3))
        Tb = synth.make_T(synth._random_rotation(rng), rng.uniform(-0.5, 0.5, size=3))
        for g1, g2 in ((ga, gb), (ga, gc), (gc, gc), (gb, gc)):
            d, db = o.geom_distance(g1, Ta, g2, Tb), o.geom_distance_brute(g1, Ta, g2, Tb)
            assert d == pytest.approx(db, abs=1e-13)
            assert o.geom_collides(g1, Ta, g2, Tb) == (db <= 0)


def test_point_cloud_distance_against_ckdtree():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(6)
    P, Qp = rng.uniform(-1, 1, size=(3000, 3)), rng.uniform(-1, 1, size=(200, 3)) + [2.5, 0, 0]
    w = WorldSpec()
    gp, gq = w.add_geom(GeomSpec.cloud(P)), w.add_geom(GeomSpec.cloud(Qp))
    w.robot = synth.make_planar_nR(w, 1)
    o = OracleWorld(w)
    T = synth.make_T(synth._random_rotation(rng), [0.1, -0.2, 0.3])
    Qw = synth.transform_points(T, Qp)
    want = cKDTree(P).query(Qw)[0].min()
    assert o.geom_distance(gp, I12, gq, T) == pytest.approx(want, abs=1e-13)


# ------------------------------------------------------------------------------------------ FK
def test_planar_nR_closed_form_fk():
    w = WorldSpec()
    n, Llen = 5, 0.7
    w.robot = synth.make_planar_nR(w, n, Llen)
    o = OracleWorld(w)
    rng = np.random.default_rng(2)
    for _ in range(20):
        q = rng.uniform(0, 6.28, size=n)
        T = o.fk(q)
        x = z = 0.0
        th = 0.0
        for i in range(n):
            if i > 0:
                x += Llen * math.cos(th)
                z += -Llen * math.sin(th)        # rotation about +y takes +x towards -z
            th += q[i]
            R = T[i, :9].reshape(3, 3)
            assert np.allclose(T[i, 9:], [x, 0, z], atol=1e-12)
            assert np.allclose(R, synth.rot_axis_angle([0, 1, 0], th), atol=1e-12)


def test_fk_prismatic_and_branching():
    w = WorldSpec()
    r = synth.make_planar_nR(w, 3)
    r.linktype = np.array([0, 1, 0], dtype=np.uint8)
    r.parents = np.array([-1, 0, 0], dtype=np.int32)
    r.axis[1] = [0, 0, 1]
    w.robot = r
    o = OracleWorld(w)
    T = o.fk([math.pi / 2, 0.25, 0.1])
    # link 1: prismatic along its local z after the parent's rotation about y and the 1 m offset along x
    assert np.allclose(T[1, 9:], synth.rot_axis_angle([0, 1, 0], math.pi / 2) @ [1, 0, 0.25], atol=1e-12)
    assert np.allclose(T[1, :9].reshape(3, 3), synth.rot_axis_angle([0, 1, 0], math.pi / 2), atol=1e-12)
    assert np.allclose(T[2, :9].reshape(3, 3), synth.rot_axis_angle([0, 1, 0], math.pi / 2 + 0.1), atol=1e-12)


# ------------------------------------------------------------------------------------------ limits, mask, feasibility
def test_joint_and_driver_limits():
    w = synth.world_c1()
    w.robot.drivers.append(DriverSpec(links=[2, 3], scale=[1.0, -1.0], offset=[0.0, 0.0], qmin=-0.5, qmax=0.5))
    o = OracleWorld(w)
    q = np.zeros(7)
    assert o.check_joint_limits(q)
    q[1] = w.robot.qmax[1]
    assert o.check_joint_limits(q)                         # closed interval
    q[1] = np.nextafter(w.robot.qmax[1], 10)
    assert not o.check_joint_limits(q)
    q[1] = 0
    q[2], q[3] = 0.6, -0.6                                 # driver value = mean(0.6/1, -0.6/-1) = 0.6 > 0.5
    assert not o.check_joint_limits(q)
    q[2], q[3] = 0.6, 0.2                                  # mean(0.6, -0.2) = 0.2
    assert o.check_joint_limits(q)
    q[0] = 1e-12                                           # welded base: qmin = qmax = 0
    assert not o.check_joint_limits(q)


def test_default_pair_mask_quirks():
    """WorldPlannerSettings::InitializeDefault (Cpp/Planning/PlannerSettings.cpp:16-41)"""
    w = synth.world_c1()
    o = OracleWorld(w)
    m = o.pair_mask()
    n = w.num_ids()
    assert m.shape == (n, n)
    rid, base = w.robot_id(), w.robot_link_id(0)
    assert m[rid, rid] == 1                                               # robot can self collide
    assert all(m[i, i] == 0 for i in range(n) if i != rid)
    assert not m[base:base + 7, rid].any() and not m[rid, base:base + 7].any()
    assert m[base + 0, 0] == 0 and m[0, base + 0] == 0                     # root link vs terrain
    assert m[base + 1, 0] == 1                                            # other links vs terrain
    assert m[base + 1, 1] == 1 and m[1, base + 1] == 1                     # link vs rigid object
    L = 7
    for i in range(L):
        for j in range(L):
            want = 1 if (i < j and j != i + 1) else 0                      # upper triangular, parent-child excluded
            assert m[base + i, base + j] == want
    # 15 default self pairs for the 7-link chain: C(7,2) - 6
    assert m[base:base + L, base:base + L].sum() == 15


def test_self_collision_edits_and_empty_links():
    w = synth.world_c3()
    r = w.robot
    assert r.link_geom.count(-1) == 2                                      # the two welded mounting frames carry no geometry
    r.self_collision_edits += [(1, 3, False), (1, 2, True)]
    o = OracleWorld(w)
    m = o.pair_mask()
    base = w.robot_link_id(0)
    assert m[base + 1, base + 3] == 0 and m[base + 1, base + 2] == 0       # (1,2): link 2 has no geometry -> cannot be enabled
    mount = [i for i, g in enumerate(r.link_geom) if g < 0]
    assert not m[base + mount[0], :].any()


def test_bvh_feasibility_equals_brute_force():
    """full IsFeasible (mask + AABB broad phase + BVH descent) against all-pairs brute force on a coarse world"""
    rng = np.random.default_rng(9)
    w = WorldSpec()
    v, t = synth.box_mesh([-2, -2, -0.6], [2, 2, -0.5], div=2)
    w.terrains.append(w.add_geom(GeomSpec.mesh(v, t)))
    for _ in range(6):
        d = rng.uniform(0.1, 0.4, size=3)
        v, t = synth.box_mesh(-d, d, div=2)
        w.objects.append((w.add_geom(GeomSpec.mesh(v, t)), synth.make_T(synth._random_rotation(rng), rng.uniform(-2.5, 2.5, size=3) * [1, 0.3, 1])))
    w.robot = synth.make_planar_nR(w, 5, 0.8)
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 400, 77)
    got = np.array([o.feasible(q) for q in Q])
    want = np.array([o.feasible_brute(q) for q in Q])
    assert (got == want).all()
    assert 0.1 < got.mean() < 0.9
    # and a handful of configurations of the full-resolution C1 world
    w1 = synth.world_c1()
    o1 = OracleWorld(w1)
    for q in synth.sample_configs(w1.robot, 3, 79):
        assert o1.feasible(q) == o1.feasible_brute(q)


def test_traversal_counts_definition():
    """the roofline's algorithmic bytes use these counters: they must be deterministic and non-trivial"""
    w = synth.world_c1()
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 200, 78)
    out, pairs, cnt = o.feasible_batch(Q, nthreads=1, want_pairs=True, want_counts=True)
    out2, cnt2 = o.feasible_batch(Q, nthreads=0, want_counts=True)
    assert (out == out2).all() and (cnt == cnt2).all()
    assert cnt["n_box"].min() >= 7 and cnt["n_node"].sum() > 0 and cnt["n_pt"].sum() == 0
    assert ((pairs[:, 0] >= 0) == ((out == 0) & np.array([o.check_joint_limits(q) for q in Q]))).all()


# ------------------------------------------------------------------------------------------ edges
def _py_edge(o, a, b, eps):
    length = o.cspace_distance(a, b)
    segs, n = 1, 0
    while length > eps:
        segs *= 2
        length *= 0.5
        for k in range(1, segs, 2):
            n += 1
            if not o.feasible(o.interpolate(a, b, k / segs)):
                return False, n
    return True, n


def test_edge_checker_order_and_counts():
    w = synth.world_c1()
    o = OracleWorld(w)
    A, B = synth.sample_edges(w.robot, lambda Q: o.feasible_batch(Q), 40, 3)
    vis, n = o.edges_visible_batch(A, B, eps=0.05)
    for i in range(len(A)):
        v2, n2 = _py_edge(o, A[i], B[i], 0.05)
        assert bool(vis[i]) == v2 and n[i] == n2
    # a free edge of length l costs 2^ceil(log2(l/eps)) - 1 checks
    free = np.nonzero(vis)[0]
    assert len(free) > 0
    for i in free[:5]:
        l = o.cspace_distance(A[i], B[i])
        assert n[i] == 2 ** max(0, math.ceil(math.log2(l / 0.05))) - 1
    # zero-length edge: visible with no checks; endpoints are never re-checked
    v, k = o.edge_visible(A[0], A[0], 0.01)
    assert v and k == 0


def test_metric_and_interpolation_spin_joint():
    w = WorldSpec()
    r = synth.make_planar_nR(w, 2)
    r.joint_type = np.array([JOINT_SPIN, JOINT_NORMAL], dtype=np.uint8)
    w.robot = r
    o = OracleWorld(w)
    a, b = np.array([0.1, 1.0]), np.array([2 * math.pi - 0.1, 2.0])
    assert o.cspace_distance(a, b) == pytest.approx(math.hypot(0.2, 1.0))      # short way round for the spin joint
    m = o.interpolate(a, b, 0.5)
    assert m[1] == pytest.approx(1.5)
    assert min(m[0], 2 * math.pi - m[0]) == pytest.approx(0.0, abs=1e-12)
    assert o.cspace_distance(a, b, weights=[4.0, 1.0]) == pytest.approx(math.sqrt(4 * 0.04 + 1.0))


def test_robot_distance_lower_bound():
    w = synth.world_c1()
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 30, 5)
    for q in Q:
        d, pair = o.distance(q, upper_bound=0.4, include_self=False)
        assert 0 <= d <= 0.4
        feas_env = d > 0
        if not feas_env:
            assert not o.feasible(q) or not o.check_joint_limits(q)
        if d < 0.4:
            assert pair[0] >= w.robot_link_id(0) and 0 <= pair[1] < w.robot_id()


def test_robot_distance_with_margins_is_the_exact_minimum():
    """inside the margins distances go negative; the robot-level minimum must still be the minimum over ALL enabled pairs
    (an AABB distance of 0 bounds nothing there)"""
    w = synth.world_c5(n_points=4000, n_obstacles=20, margin=0.02)
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 300, 52)
    d, _ = o.distance_batch(Q, upper_bound=0.5, include_self=False)
    cloud = w.terrains[0]
    neg = np.nonzero(d < 0)[0][:12]
    assert len(neg) > 3
    for i in neg:
        T = o.fk(Q[i])
        best = min(o.geom_distance_brute(g, T[j], cloud, I12) for j, g in enumerate(w.robot.link_geom) if j > 0)
        assert d[i] == pytest.approx(best, abs=1e-13)


def test_metric_and_interpolation_floating_and_ball_joints():
    """Floating / BallAndSocket joints (reference Cpp/Modeling/Interpolate.cpp:16-52,229-278): Euler ZYX triplets are
    compared by geodesic angle and interpolated along the SO(3) geodesic.  Second method: scipy Rotation / Slerp."""
    from scipy.spatial.transform import Rotation as Rot, Slerp
    w = synth.world_floating()
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 120, 3)
    rng = np.random.default_rng(0)
    T = o.fk_batch(Q[:8])
    for i in range(8):     # the three revolute links of the floating joint compose to Rz(a) Ry(b) Rx(c)
        np.testing.assert_allclose(T[i, 5, :9].reshape(3, 3), Rot.from_euler("ZYX", Q[i, 3:6]).as_matrix(), atol=1e-14)
    for i in range(60):
        a, b = Q[2 * i], Q[2 * i + 1]
        Ra, Rb = Rot.from_euler("ZYX", a[3:6]), Rot.from_euler("ZYX", b[3:6])
        Wa, Wb = Rot.from_euler("ZYX", a[7:10]), Rot.from_euler("ZYX", b[7:10])
        want = math.sqrt(((a[:3] - b[:3]) ** 2).sum() + (Ra * Rb.inv()).magnitude() ** 2 + (a[6] - b[6]) ** 2 + (Wa * Wb.inv()).magnitude() ** 2)
        assert abs(o.cspace_distance(a, b) - want) < 1e-9
        u = rng.uniform()
        m = o.interpolate(a, b, u)
        Rm, Wm = Slerp([0, 1], Rot.concatenate([Ra, Rb]))(u), Slerp([0, 1], Rot.concatenate([Wa, Wb]))(u)
        assert (Rot.from_euler("ZYX", m[3:6]) * Rm.inv()).magnitude() < 1e-7
        assert (Rot.from_euler("ZYX", m[7:10]) * Wm.inv()).magnitude() < 1e-7
        np.testing.assert_allclose(m[:3], a[:3] * (1 - u) + b[:3] * u, atol=1e-14)
        assert abs(m[6] - (a[6] * (1 - u) + b[6] * u)) < 1e-14
    # end points are reproduced as rotations, and a half-turn apart pair still interpolates on the geodesic
    a = np.zeros(10); b = np.zeros(10); b[3] = math.pi - 1e-9
    m = o.interpolate(a, b, 0.5)
    assert abs(abs(m[3]) - math.pi / 2) < 1e-6 and abs(m[4]) < 1e-6 and abs(m[5]) < 1e-6


def test_multi_link_joint_layout_is_validated():
    """The reference asserts the link layout of Floating / BallAndSocket joints (Interpolate.cpp:24-26,231-236)."""
    w = synth.world_floating()
    w.robot.joint_base = np.array([0, 5, 6], dtype=np.int32)      # floating joint would drive only 5 links
    with pytest.raises(ValueError):
        OracleWorld(w)


def test_triangle_primitive_is_a_one_triangle_mesh():
    """GeometricPrimitive "Triangle" against the unit cube of the reference's tests/objects/cube.off"""
    w = WorldSpec()
    v, t = synth.unit_cube()
    gc = w.add_geom(GeomSpec.mesh(v, t))
    gt = w.add_geom(GeomSpec.triangle([0, 0, 0], [1, 0, 0], [0, 1, 0]))
    w.robot = synth.make_planar_nR(w, 1)
    o = OracleWorld(w)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    T = I.copy(); T[9:] = [0.25, 0.25, 1.0 - 1e-3]
    assert o.geom_collides(gt, T, gc, I)
    T[11] = 1.5
    assert not o.geom_collides(gt, T, gc, I)
    assert abs(o.geom_distance(gt, T, gc, I) - 0.5) < 1e-12
    assert o.geom_within_distance(gt, T, gc, I, 0.5 + 1e-9) and not o.geom_within_distance(gt, T, gc, I, 0.5 - 1e-9)


def test_solid_box_primitive_semantics():
    """Box / AABB primitives are solid (GeometricPrimitive3D Box3D / AABB3D; the common primitives of
    Cpp/docs/Manual-Geometry.md:241-250): an element inside the box collides although no surfaces meet; outside, the
    distance is the distance to the box surface.  Second method: closed-form point-to-box distance in numpy."""
    w = WorldSpec()
    v, t = synth.unit_cube()
    gc = w.add_geom(GeomSpec.mesh(v * 0.2, t))
    gb = w.add_geom(GeomSpec.aabb([0, 0, 0], [1, 1, 1]))
    gs = w.add_geom(GeomSpec.sphere([0, 0, 0], 0.1))
    R = synth.rot_axis_angle([1, 2, 3], 0.7)
    go = w.add_geom(GeomSpec.box([0.1, -0.2, 0.3], R, [0.5, 0.25, 0.125]))
    gp = w.add_geom(GeomSpec.point([0, 0, 0]))
    w.robot = synth.make_planar_nR(w, 1)
    o = OracleWorld(w)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    T = I.copy(); T[9:] = [0.4, 0.4, 0.4]                       # the small cube is wholly inside the unit box
    assert o.geom_collides(gc, T, gb, I) and o.geom_collides(gb, I, gc, T) and o.geom_distance(gc, T, gb, I) == 0.0
    T[9:] = [1.5, 0.4, 0.4]
    assert not o.geom_collides(gc, T, gb, I) and abs(o.geom_distance(gc, T, gb, I) - 0.5) < 1e-12
    assert abs(o.geom_distance_brute(gc, T, gb, I) - 0.5) < 1e-12
    T[9:] = [0.5, 0.5, 0.5]
    assert o.geom_collides(gs, T, gb, I) and abs(o.geom_distance(gs, T, gb, I) + 0.1) < 1e-12     # centre inside: -radius
    T[9:] = [1.3, 0.5, 0.5]
    assert not o.geom_collides(gs, T, gb, I) and abs(o.geom_distance(gs, T, gb, I) - 0.2) < 1e-12
    assert o.geom_within_distance(gs, T, gb, I, 0.2 + 1e-9) and not o.geom_within_distance(gs, T, gb, I, 0.2 - 1e-9)
    # oriented box vs points at random poses of both
    rng = np.random.default_rng(3)
    c, h = np.array([0.1, -0.2, 0.3]), np.array([0.5, 0.25, 0.125])
    for _ in range(200):
        Tb = np.concatenate([synth._random_rotation(rng).reshape(-1), rng.uniform(-0.5, 0.5, size=3)])
        Tp = I.copy(); Tp[9:] = rng.uniform(-1.2, 1.2, size=3)
        pl = Tb[:9].reshape(3, 3).T @ (Tp[9:] - Tb[9:])            # the point in the box geometry's local frame
        q = R.T @ (pl - c)
        want = float(np.linalg.norm(np.maximum(np.abs(q) - h, 0.0)))
        assert abs(o.geom_distance(go, Tb, gp, Tp) - want) < 1e-12
        assert o.geom_collides(go, Tb, gp, Tp) == (want == 0.0)


def test_world_with_solid_boxes_bvh_equals_brute_force():
    w = synth.world_boxes(n_boxes=10, n_blobs=1)
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 300, 17)
    f = o.feasible_batch(Q)
    assert 0.05 < f.mean() < 0.95
    for i in range(6):                                           # the all-pairs brute force costs seconds per configuration
        assert bool(f[i]) == o.feasible_brute(Q[i])


def test_zero_area_triangles_are_segments_not_planes():
    """a triangle with repeated or collinear vertices makes every orientation test against it vanish; it must still behave as the
    segment it is (an earlier version of the oracle sent it down the coplanar path and reported hits up to millimetres away --
    found by the GPU parity test on slivers)"""
    T = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    def sliver(p, q, collinear=False):
        p, q = np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)
        return np.array([p, q, p + 0.37 * (q - p) if collinear else q])
    for col in (False, True):
        assert not ko.tri_tri_intersect(sliver([0.1, 0.1, 1e-3], [0.4, 0.3, 1e-3], col), T)      # hovering 1 mm above, parallel
        assert not ko.tri_tri_intersect(T, sliver([0.1, 0.1, 1e-3], [0.4, 0.3, 2e-3], col))
        assert ko.tri_tri_intersect(sliver([0.2, 0.2, -0.5], [0.2, 0.2, 0.5], col), T)           # piercing the interior
        assert ko.tri_tri_intersect(T, sliver([0.2, 0.2, -0.5], [0.2, 0.2, 0.5], col))
        assert not ko.tri_tri_intersect(sliver([0.8, 0.8, -0.5], [0.8, 0.8, 0.5], col), T)       # piercing the plane outside
        assert ko.tri_tri_intersect(sliver([-0.5, 0.2, 0.0], [0.5, 0.2, 0.0], col), T)           # lying in the plane, crossing
        assert not ko.tri_tri_intersect(sliver([-0.5, -0.2, 0.0], [0.5, -0.2, 0.0], col), T)     # in the plane, beside
        assert abs(ko.tri_tri_distance(sliver([0.1, 0.1, 1e-3], [0.4, 0.3, 1e-3], col), T) - 1e-3) < 1e-15
    # two slivers: segments in space
    assert ko.tri_tri_intersect(sliver([0, 0, 0], [1, 1, 0]), sliver([0, 1, 0], [1, 0, 0]))
    assert not ko.tri_tri_intersect(sliver([0, 0, 0], [1, 1, 0]), sliver([0, 1, 0.01], [1, 0, 0.01]))


def _np_seg_seg_dist(p1, q1, p2, q2):
    """second method: the squared distance is a convex quadratic in (s, t) on the unit square -- its minimum is the interior
    stationary point if that lies inside, else the best of the four edges (each a clamped 1-D projection)"""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    def f(s, t):
        v = r + s * d1 - t * d2
        return float(v @ v)
    best = min(f(0, 0), f(0, 1), f(1, 0), f(1, 1))
    A = np.array([[d1 @ d1, -(d1 @ d2)], [-(d1 @ d2), d2 @ d2]]); b = -np.array([d1 @ r, -(d2 @ r)])
    if abs(np.linalg.det(A)) > 1e-14 * A[0, 0] * A[1, 1]:
        s, t = np.linalg.solve(A, b)
        if 0 <= s <= 1 and 0 <= t <= 1:
            best = min(best, f(s, t))
    for s in (0.0, 1.0):
        t = min(max(((r + s * d1) @ d2) / (d2 @ d2), 0.0), 1.0); best = min(best, f(s, t))
    for t in (0.0, 1.0):
        s = min(max(-((r - t * d2) @ d1) / (d1 @ d1), 0.0), 1.0); best = min(best, f(s, t))
    return math.sqrt(best)


def test_segment_primitive_semantics():
    """GeometricPrimitive "Segment" (Segment3D; Cpp/docs/Manual-Geometry.md:22): closed-form answers against the unit cube of
    tests/objects/cube.off, a sphere and other segments; margins widen it into a capsule"""
    rng = np.random.default_rng(5)
    w = WorldSpec()
    v, t = synth.unit_cube()
    gc = w.add_geom(GeomSpec.mesh(v, t))
    gs = w.add_geom(GeomSpec.segment([0.5, 0.5, 1.25], [0.5, 0.5, 2.0]))
    gsph = w.add_geom(GeomSpec.sphere([0, 0, 0], 0.1))
    gcap = w.add_geom(GeomSpec.segment([0, 0, 0], [1, 0, 0], margin=0.05))
    segs = [(rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)) for _ in range(40)]
    gr = [w.add_geom(GeomSpec.segment(a, b)) for a, b in segs]
    w.robot = synth.make_planar_nR(w, 1)
    o = OracleWorld(w)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    assert abs(o.geom_distance(gs, I, gc, I) - 0.25) < 1e-15 and not o.geom_collides(gs, I, gc, I)
    T = I.copy(); T[11] = -0.5                                        # now it pierces the top face
    assert o.geom_collides(gs, T, gc, I) and o.geom_distance(gs, T, gc, I) == 0.0
    T[11] = -1.125                                                    # wholly inside the cube's surface: a mesh is not a solid
    assert not o.geom_collides(gs, T, gc, I) and abs(o.geom_distance(gs, T, gc, I) - 0.125) < 1e-15
    # sphere beside the capsule's axis: distance = gap - radius - margin, signed
    T = I.copy(); T[9:] = [0.5, 0.3, 0.0]
    assert abs(o.geom_distance(gsph, T, gcap, I) - (0.3 - 0.1 - 0.05)) < 1e-15
    T[10] = 0.12
    assert o.geom_collides(gsph, T, gcap, I) and abs(o.geom_distance(gsph, T, gcap, I) - (0.12 - 0.15)) < 1e-15
    with pytest.raises(ValueError):
        w2 = WorldSpec(); w2.add_geom(GeomSpec.segment([1, 2, 3], [1, 2, 3])); w2.robot = synth.make_planar_nR(w2, 1); OracleWorld(w2)
    # segment pairs at random poses against the numpy closed form
    for k in range(0, 40, 2):
        Ta, Tb = synth.make_T(synth._random_rotation(rng), rng.uniform(-0.5, 0.5, 3)), synth.make_T(synth._random_rotation(rng), rng.uniform(-0.5, 0.5, 3))
        Ra, ta, Rb, tb = Ta[:9].reshape(3, 3), Ta[9:], Tb[:9].reshape(3, 3), Tb[9:]
        want = _np_seg_seg_dist(Ra @ segs[k][0] + ta, Ra @ segs[k][1] + ta, Rb @ segs[k + 1][0] + tb, Rb @ segs[k + 1][1] + tb)
        assert abs(o.geom_distance(gr[k], Ta, gr[k + 1], Tb) - want) < 1e-12
        assert o.geom_within_distance(gr[k], Ta, gr[k + 1], Tb, want + 1e-9) and not o.geom_within_distance(gr[k], Ta, gr[k + 1], Tb, want - 1e-9)


def test_contact_depth_known_answers(cubes):
    """ko_penetration / ko_geom_penetration: the colliding side of the two-sided 1e-6 m band (tests/parity.py)"""
    from oracle.oracle import tri_tri_depth
    o, ga, gb, gm, gs = cubes
    # two triangles crossing like a plus sign: A in the plane z = 0, B in the plane y = 0 dipping 0.01 below A
    A = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], float)
    B = np.array([[-0.2, 0, -0.01], [0.2, 0, -0.01], [0, 0, 1.0]], float)
    assert ko.tri_tri_intersect(A, B)
    assert ko.tri_tri_depth(A, B) == pytest.approx(0.01, abs=1e-15)          # lift B by 0.01 along z and they separate
    # unit cubes overlapping by delta along x: depth = delta (faces x = 1 of A and x = 0 of B cross the other cube's side faces)
    for delta in (1e-3, 1e-7):
        assert o.geom_penetration(ga, I12, gb, T_at(1.0 - delta, 0.25, 0.25)) == pytest.approx(delta, rel=1e-6)
    assert o.geom_penetration(ga, I12, gb, T_at(1.5)) == -1.0            # nothing in contact
    # margins: threshold 0.1, gap 0.05 -> depth 0.05; with tol the threshold grows
    assert o.geom_penetration(ga, I12, gm, T_at(1.05)) == pytest.approx(0.05, abs=1e-12)
    assert o.geom_penetration(ga, I12, gb, T_at(1.05), 0.08) == pytest.approx(0.03, abs=1e-12)
    # sphere of radius 0.25 centred 0.2 outside the face x = 1: depth 0.05
    assert o.geom_penetration(ga, I12, gs, T_at(1.2, 0.5, 0.5)) == pytest.approx(0.05, abs=1e-12)


def test_contact_depth_of_a_robot_configuration():
    w = synth.world_c1()
    o = OracleWorld(w)
    Q = synth.sample_configs(w.robot, 400, 77)
    feas = o.feasible_batch(Q)
    for i in range(len(Q)):
        p = o.penetration(Q[i])
        if feas[i]:
            assert p == -1.0
        elif o.check_joint_limits(Q[i]):
            assert p >= 0.0


def test_fast_baseline_variant_agrees_with_the_checker():
    """libkb_oracle_fast.so (SAH tree, -march=native, FMA allowed) is the CPU arm bench.py times; it must answer like the checker"""
    for w, n in ((synth.world_c1(), 3000), (synth.world_c2(2, n_obstacles=30), 3000), (synth.world_c3(), 1500)):
        o, f = OracleWorld(w), OracleWorld(w, variant="fast")
        Q = synth.sample_configs(w.robot, n, 5)
        a, b = o.feasible_batch(Q), f.feasible_batch(Q)
        bad = np.nonzero(a != b)[0]
        for i in bad:                                  # FMA contraction may flip a contact that is exactly on the boundary
            assert abs(o.distance(Q[i], upper_bound=1.0, include_self=True)[0]) <= 1e-9 or o.penetration(Q[i]) <= 1e-9
        assert len(bad) <= 2
        da, _ = o.distance_batch(Q[:200], upper_bound=0.5)
        db, _ = f.distance_batch(Q[:200], upper_bound=0.5)
        np.testing.assert_allclose(da, db, rtol=1e-9, atol=1e-12)


# ------------------------------------------------------------------------------------------ ray casting
def test_raycast_known_answers(cubes):
    """Geometry3D::rayCast_ext on the unit cube [0,1]^3 and a sphere: analytic distances, margin taken off a mesh's distance, a moved
    geometry, rays that miss, start inside or point away; the hierarchy against every-element brute force"""
    o, ga, gb, gm, gs = cubes
    hit, d, el = o.geom_raycast(ga, I12, [0.25, 0.5, 3.0], [0, 0, -1])
    assert hit and d == pytest.approx(2.0, abs=1e-15) and el >= 0
    hit, d, _ = o.geom_raycast(ga, I12, [0.25, 0.5, 3.0], [0, 0, -7.5])                      # direction of any length: dist is a length
    assert hit and d == pytest.approx(2.0, abs=1e-15)
    assert not o.geom_raycast(ga, I12, [0.25, 0.5, 3.0], [0, 0, 1])[0]                       # pointing away
    assert not o.geom_raycast(ga, I12, [1.5, 0.5, 3.0], [0, 0, -1])[0]                       # passes beside the cube
    hit, d, _ = o.geom_raycast(ga, I12, [0.5, 0.5, 0.5], [1, 0, 0])                          # from inside: the far face (surface mesh, two-sided)
    assert hit and d == pytest.approx(0.5, abs=1e-15)
    hit, d, _ = o.geom_raycast(gm, I12, [0.25, 0.5, 3.0], [0, 0, -1])                        # margin 0.1 comes off the distance
    assert hit and d == pytest.approx(1.9, abs=1e-15)
    hit, d, _ = o.geom_raycast(ga, T_at(0.0, 0.0, -2.0), [0.25, 0.5, 3.0], [0, 0, -1])       # cube moved down by 2
    assert hit and d == pytest.approx(4.0, abs=1e-15)
    s = 1 / math.sqrt(3)
    hit, d, _ = o.geom_raycast(gs, I12, [2.0, 2.0, 2.0], [-s, -s, -s])                       # sphere of radius 0.25 at the origin
    assert hit and d == pytest.approx(2 * math.sqrt(3) - 0.25, abs=1e-14)
    assert not o.geom_raycast(gs, I12, [2.0, 2.0, 2.0], [-1, 0, 0])[0]
    hit, d, _ = o.geom_raycast(gs, I12, [0.1, 0.0, 0.0], [1, 0, 0])                          # source inside the sphere
    assert hit and d == 0.0
    assert not o.geom_raycast(ga, I12, [0.25, 0.5, 3.0], [0, 0, 0])[0]                       # no direction
    rng = np.random.default_rng(2)
    w = synth.world_c1()
    ow = OracleWorld(w)
    for g in range(min(4, len(w.geoms))):
        lo, hi = ow.geom_aabb(g, I12)
        for _ in range(200):
            src = 0.5 * (lo + hi) + rng.normal(size=3) * np.linalg.norm(hi - lo)
            dirn = rng.uniform(lo, hi) - src
            a, b = ow.geom_raycast(g, I12, src, dirn), ow.geom_raycast(g, I12, src, dirn, brute=True)
            assert a[0] == b[0] and (not a[0] or (abs(a[1] - b[1]) < 1e-12 and a[2] == b[2]))


def test_world_raycast_order_ignore_and_robot():
    """WorldModel::RayCast (World.cpp:465-516): nearest body wins, ids follow the world numbering; RayCastIgnore skips ids; the robot is
    cast at the given configuration"""
    w = WorldSpec()
    v, t = synth.unit_cube()
    g = w.add_geom(GeomSpec.mesh(v, t))
    w.terrains.append(g)                                             # id 0: cube at [0,1]^3
    w.objects.append((g, T_at(3.0)))                                 # id 1: cube at x in [3,4]
    w.objects.append((g, T_at(6.0)))                                 # id 2: cube at x in [6,7]
    w.robot = synth.make_planar_nR(w, 2)
    o = OracleWorld(w)
    rays = np.array([[10.0, 0.5, 0.5, -1, 0, 0], [-5.0, 0.5, 0.5, 1, 0, 0], [4.5, 0.5, 0.5, 1, 0, 0], [4.5, 0.5, 5.0, 0, 0, 1]])
    ids, dist, _ = o.raycast_batch(None, rays)
    assert list(ids) == [2, 0, 2, -1]
    np.testing.assert_allclose(dist[:3], [3.0, 5.0, 1.5], atol=1e-14)
    assert np.isinf(dist[3])
    ig = np.zeros(o.num_ids(), dtype=np.uint8)
    ig[2] = 1
    ids, dist, _ = o.raycast_batch(None, rays, ig)
    assert list(ids) == [1, 0, -1, -1] and dist[0] == pytest.approx(6.0, abs=1e-14)
    q = np.zeros(w.robot.L)
    ids_r, dist_r, _ = o.raycast_batch(q, rays)
    T = o.fk(q)
    for j in range(w.robot.L):                                       # every link geometry is hit by a ray aimed at its box centre from above
        gj = w.robot.link_geom[j]
        if gj < 0:
            continue
        lo, hi = o.geom_aabb(gj, T[j])
        c = 0.5 * (lo + hi)
        i1, d1, _ = o.raycast_batch(q, [[c[0], c[1], hi[2] + 1.0, 0, 0, -1]])
        if i1[0] == w.robot_link_id(j):
            assert d1[0] <= 1.0 + (hi[2] - lo[2]) + 1e-12
            break
    else:
        pytest.fail("no link was hit")


def test_raycast_is_watertight_on_shared_edges_and_vertices(cubes):
    """rays aimed exactly at vertices, edge points and face diagonals of the unit cube (each face is two triangles) never slip between
    the triangles: the edge functions of the watertight test are exact in sign and antisymmetric in an edge's two vertices"""
    o, ga, gb, gm, gs = cubes
    targets = [(x, y) for x in (0.0, 0.25, 0.5, 1.0) for y in (0.0, 0.25, 0.5, 1.0)] + [(t, t) for t in np.linspace(0, 1, 23)] + [(t, 1 - t) for t in np.linspace(0, 1, 23)]
    rng = np.random.default_rng(4)
    for (x, y) in targets:
        hit, d, _ = o.geom_raycast(ga, I12, [x, y, 3.0], [0, 0, -1])                     # straight down onto the top face z = 1
        assert hit and d == pytest.approx(2.0, abs=1e-15), (x, y)
        if not (0.0 < x < 1.0 and 0.0 < y < 1.0):
            continue                                                                   # the cube's own edges are silhouettes: grazing them may miss
        for _ in range(4):                                                             # from a random source through the same point of the face diagonal
            src = np.array([x, y, 1.0]) + rng.uniform([-2, -2, 0.5], [2, 2, 3.0])
            hit, d, _ = o.geom_raycast(ga, I12, src, np.array([x, y, 1.0]) - src)
            assert hit and d <= np.linalg.norm(np.array([x, y, 1.0]) - src) + 1e-12, (x, y)
