"""Ray casting against the CPU oracle: WorldModel::RayCast / RayCastIgnore semantics (World.cpp:465-588) through
kb_raycast_batch, Geometry3D::rayCast_ext through kb_geom_raycast_batch.  Hit / miss and the body hit must be equal; distances
within 1e-9 relative (both sides evaluate the same fp64 ray / triangle statement; only the hierarchy differs); the reported element
must reproduce the distance when it is cast alone (two triangles sharing the edge a ray crosses tie)."""
import numpy as np
import pytest

from klampt_b200 import synth

pytestmark = pytest.mark.gpu


def camera_rays(eye, target, up, fov_deg, w, h):
    """pinhole camera rays the way the camera sensor's fallback builds them (VisualSensors.cpp:430-450): unit directions"""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    fx = 0.5 * w / np.tan(0.5 * np.radians(fov_deg))
    ii, jj = np.meshgrid(np.arange(w) - 0.5 * (w - 1), 0.5 * (h - 1) - np.arange(h))
    d = f[None, None, :] + (ii / fx)[..., None] * r + (jj / fx)[..., None] * u
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return np.concatenate([np.broadcast_to(eye, d.shape), d], axis=-1).reshape(-1, 6)


def check_against_oracle(eng, orc, q, rays, ignore=None, rtol=1e-9):
    ids, dist, elem = eng.raycast_batch(q, rays, ignore)
    ig = None
    if ignore is not None:
        ig = np.zeros(orc.num_ids(), dtype=np.uint8)
        ig[list(ignore)] = 1
    oid, odist, oelem = orc.raycast_batch(q, rays, ig)
    assert np.array_equal(ids >= 0, oid >= 0), "hit / miss differs on %d rays" % int(((ids >= 0) != (oid >= 0)).sum())
    hit = ids >= 0
    assert np.all(np.isinf(dist[~hit])) and np.all(elem[~hit] == -1)
    np.testing.assert_allclose(dist[hit], odist[hit], rtol=rtol, atol=1e-12)
    assert np.array_equal(ids, oid)
    return ids, dist, elem, oelem


def test_camera_image_of_the_c2_world(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2(2, n_obstacles=40)
    eng, orc = Engine(w), OracleWorld(w)
    q = synth.sample_configs(w.robot, 1, 3)[0]
    rays = camera_rays((2.6, 0.4, 1.4), (0, 0, 0.5), (0, 0, 1), 70, 160, 120)
    ids, dist, elem, oelem = check_against_oracle(eng, orc, q, rays)
    assert 0.3 < (ids >= 0).mean() < 1.0
    assert len(np.unique(ids[ids >= 0])) > 5                       # links and several obstacles are in view
    agree = (elem == oelem)[ids >= 0].mean()
    assert agree > 0.999                                           # ties across a shared edge aside
    # the robot left out, and single links ignored (RayCastIgnore, the laser sensor's use)
    ids2, _, _, _ = check_against_oracle(eng, orc, None, rays)
    assert not np.any(ids2 >= w.robot_id())
    link_ids = [w.robot_link_id(j) for j in range(w.robot.L)]
    ids3, _, _, _ = check_against_oracle(eng, orc, q, rays, ignore=link_ids[2:5])
    assert not np.isin(ids3, link_ids[2:5]).any()
    # idempotent, independent of the batch split
    a = eng.raycast_batch(q, rays[:5000])
    b = eng.raycast_batch(q, rays)
    assert all(np.array_equal(x, y[:5000]) for x, y in zip(a, b))


def test_random_rays_and_degenerate_input(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2(5, n_obstacles=25)
    eng, orc = Engine(w), OracleWorld(w)
    q = synth.sample_configs(w.robot, 1, 8)[0]
    rng = np.random.default_rng(0)
    N = 20000
    src = rng.uniform([-3, -3, -0.5], [3, 3, 3], (N, 3))
    tgt = rng.uniform([-1, -1, 0], [1, 1, 1.5], (N, 3))
    rays = np.hstack([src, (tgt - src) * rng.uniform(0.1, 5.0, (N, 1))])          # directions of any length
    rays[:50, 3:] = np.eye(3)[rng.integers(0, 3, 50)] * rng.choice([-1.0, 1.0], (50, 1))     # axis-parallel: zero direction components
    check_against_oracle(eng, orc, q, rays)
    # bodies the collision mask has switched off are still seen by rays (WorldModel::RayCast knows no mask): they are not part of a
    # merged group and are cast one by one in their own frames under the top-level hierarchy
    import copy
    w2 = copy.deepcopy(w)
    m = orc.pair_mask().copy()
    off = [w2.rigid_object_id(k) for k in (0, 3, 4, 11, 17)]
    m[off, :] = 0
    m[:, off] = 0
    w2.pair_mask = m
    eng2, orc2 = Engine(w2), OracleWorld(w2)
    ids2, _, _, _ = check_against_oracle(eng2, orc2, q, rays)
    assert np.isin(ids2, off).sum() > 20
    assert np.array_equal(ids2, eng.raycast_batch(q, rays)[0])
    # fp32 rays: the fp64 answer for the rounded rays
    rf = rays.astype(np.float32)
    a32, b32 = eng.raycast_batch(q, rf), eng.raycast_batch(q, rf.astype(np.float64))
    assert all(np.array_equal(x, y) for x, y in zip(a32, b32))
    bad = rays[:4].copy()
    bad[0, 3:] = 0.0
    bad[1, 3] = np.nan
    bad[2, 0] = np.inf
    ids, dist, elem = eng.raycast_batch(q, bad)
    assert list(ids[:3]) == [-1, -1, -1] and np.all(np.isinf(dist[:3]))
    ids0, dist0, _ = eng.raycast_batch(q, np.zeros((0, 6)))
    assert ids0.shape == (0,)


def test_point_clouds_margins_and_primitives(built):
    from klampt_b200.engine import Engine
    from klampt_b200.worldspec import GeomSpec
    from oracle.oracle import OracleWorld
    w = synth.world_c5(n_points=30000, n_obstacles=12)
    eng, orc = Engine(w), OracleWorld(w)
    q = synth.sample_configs(w.robot, 1, 1)[0]
    rays = camera_rays((2.4, -0.6, 1.2), (0, 0, 0.4), (0, 0, 1), 60, 128, 96)
    ids, dist, elem, oelem = check_against_oracle(eng, orc, q, rays)
    assert (ids >= 0).mean() > 0.05
    # the same cloud through the GPU-built hierarchy
    eng2 = Engine(w, options={"cloud_builder": 1})
    ids2, dist2, _ = eng2.raycast_batch(q, rays)
    assert np.array_equal(ids2, ids)
    np.testing.assert_allclose(dist2[ids >= 0], dist[ids >= 0], rtol=1e-9, atol=1e-12)
    # boxes, spheres, a mesh with a margin
    wb = synth.world_boxes(n_boxes=8, n_blobs=2)
    eb, ob = Engine(wb), OracleWorld(wb)
    qb = synth.sample_configs(wb.robot, 1, 2)[0]
    rb = camera_rays((2.2, 0.9, 1.5), (0, 0, 0.4), (0, 0, 1), 75, 96, 96)
    check_against_oracle(eb, ob, qb, rb)


def test_geometry_raycast_ext(built):
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2(2, n_obstacles=6)
    eng, orc = Engine(w), OracleWorld(w)
    rng = np.random.default_rng(5)
    T = synth.make_T(None, (0.3, -0.2, 0.1))
    for g in (0, w.robot.link_geom[2], len(w.geoms) - 1):
        lo, hi = orc.geom_aabb(g, T)
        c, ext = 0.5 * (lo + hi), np.linalg.norm(hi - lo)
        N = 3000
        src = c + rng.normal(size=(N, 3)) * ext
        tgt = rng.uniform(lo, hi, (N, 3))
        rays = np.hstack([src, tgt - src])
        elem, dist = eng.geom_raycast_batch(g, T, rays)
        nhit = 0
        for i in range(0, N, 7):
            h, d, el = orc.geom_raycast(g, T, rays[i, :3], rays[i, 3:], brute=True)
            assert h == (elem[i] >= 0)
            if h:
                nhit += 1
                assert abs(d - dist[i]) <= 1e-9 * max(1.0, d)
        assert nhit > 20


def test_camera_and_laser_sensors_follow_the_per_ray_statement(built):
    """klampt_b200.sensing against a per-pixel restatement of CameraSensor / LaserRangeSensor::SimulateKinematic (VisualSensors.cpp:57-142,
    413-475) that calls the oracle once per ray"""
    import math
    from klampt_b200 import sensing
    from klampt_b200.engine import Engine
    from oracle.oracle import OracleWorld
    w = synth.world_c2(2, n_obstacles=30)
    eng, orc = Engine(w), OracleWorld(w)
    q = synth.sample_configs(w.robot, 1, 4)[0]
    # camera looking along -x of the world: columns of R = right, down, forward
    R = np.array([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    Tcam = synth.make_T(R, (2.5, 0.2, 0.9))
    cam = sensing.CameraSensor(xres=48, yres=32, xfov=math.radians(70), yfov=math.radians(50), zmin=0.1, zmax=3.0, Tsensor=Tcam)
    depth, ids = cam.simulate(eng, q)
    fx, fy, cx, cy = cam.viewport()
    eye, right, up, fwd = R @ np.zeros(3) + np.array([2.5, 0.2, 0.9]), R[:, 0], -R[:, 1], R[:, 2]
    want = np.empty((32, 48), dtype=np.float32)
    for j in range(32):
        for i in range(48):
            d = fwd + (i - cx) * right / fx + (cy - j) * up / fy
            src = eye + d * cam.zmin
            oid, od, _ = orc.raycast_batch(q, [np.concatenate([src, d / np.linalg.norm(d)])])
            if oid[0] >= 0:
                z = fwd @ (src + od[0] * d / np.linalg.norm(d) - eye)
                z = cam.zmax if z < cam.zmin else min(z, cam.zmax)
            else:
                z = cam.zmax
            want[j, i] = z
    np.testing.assert_allclose(depth, want, rtol=1e-6, atol=1e-6)
    assert (depth < cam.zmax).mean() > 0.2 and (ids >= 0).any()
    depth2, ids2 = cam.simulate_from_rays(eng, q)                  # host-built rays through kb_raycast_batch: the same image
    assert np.array_equal(ids2, ids)
    np.testing.assert_allclose(depth2, depth, rtol=1e-6, atol=1e-6)
    cam3 = sensing.CameraSensor(xres=40, yres=30, zmin=0.05, zmax=4.0, link=5, Tsensor=synth.make_T(None, (0.0, 0.0, 0.12)))     # camera riding on a link
    own = [w.robot_link_id(5), w.robot_link_id(6)]          # the wrist and the tool sphere the camera sits in (a source inside a sphere reads zmin exactly)
    d3, i3 = cam3.simulate(eng, q, ignore_ids=own)
    d4, i4 = cam3.simulate_from_rays(eng, q, ignore_ids=own)
    assert np.array_equal(i3, i4) and not np.isin(i3, own).any() and (i3 >= 0).mean() > 0.1
    np.testing.assert_allclose(d3, d4, rtol=1e-6, atol=1e-6)
    # laser on link 3, sweeping 90 degrees: its own link is ignored
    # (mounted at a generic angle: rays lying exactly in a symmetry plane of the link meshes graze silhouette edges, where hit or miss
    # is a matter of the last bit of FK)
    las = sensing.LaserRangeSensor(measurementCount=64, depthMinimum=0.05, depthMaximum=5.0, link=3,
                                   Tsensor=synth.make_T(synth._random_rotation(np.random.default_rng(9)), (0.01, 0.02, 0.03)))
    got = las.simulate(eng, q, link_world_id=w.robot_link_id(3))
    rays = las.rays(orc.fk(q))
    ig = np.zeros(orc.num_ids(), dtype=np.uint8)
    ig[w.robot_link_id(3)] = 1
    oid, od, _ = orc.raycast_batch(q, rays, ig)
    ref = np.where(oid >= 0, od + las.depthMinimum, np.inf)
    ref = np.where((ref <= las.depthMinimum) | (ref >= las.depthMaximum), las.depthMaximum, ref)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)


def test_python_adapters_raycast(built):
    """Geometry3D.rayCast / rayCast_ext, collide.ray_cast and WorldCollider.rayCast / rayCastRobot (reference collide.py:219-243,700-748)"""
    from klampt_b200 import collide
    from klampt_b200.robotsim import WorldModel
    spec = synth.world_c2(2, n_obstacles=8)
    world = WorldModel.from_spec(spec)
    q = synth.sample_configs(spec.robot, 1, 6)[0]
    world.robot(0).setConfig(list(q))
    wc = collide.WorldCollider(world)
    s, d = [2.5, 0.3, 1.0], [-1.0, -0.1, -0.25]
    one = wc.rayCast(s, d)
    geoms = [g for (o, g) in wc.geomList]
    loop = collide.ray_cast(geoms, s, d)
    assert (one is None) == (loop is None)
    if one is not None:
        assert one[0] is wc.geomList[loop[0]][0]
        np.testing.assert_allclose(one[1], loop[1], rtol=1e-9, atol=1e-12)
    rng = np.random.default_rng(1)
    src = np.tile([2.5, 0.3, 1.0], (300, 1))
    rays = np.hstack([src, rng.uniform([-1, -1, 0], [1, 1, 1.5], (300, 3)) - src])
    which, dist, pts = wc.rayCastBatch(rays)
    assert (which >= 0).sum() > 30
    for k in np.flatnonzero(which >= 0)[:25]:
        hit, pt = wc.geomList[which[k]][1].rayCast(rays[k, :3], rays[k, 3:])
        assert hit
        np.testing.assert_allclose(pt, pts[k], rtol=1e-9, atol=1e-12)
    links = wc.rayCastRobot(0, [0.0, 0.0, 3.0], [0.0, 0.0, -1.0])
    assert links is None or links[0].getIndex() >= 0
    el, pt = geoms[0].rayCast_ext(s, d)
    assert (el >= 0) == geoms[0].rayCast(s, d)[0]
