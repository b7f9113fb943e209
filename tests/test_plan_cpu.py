"""Roadmap planner logic on a stand-in batch space (numpy predicates): no GPU needed."""
import numpy as np
import pytest

from klampt_b200.plan import MotionPlan


class DiskSpace:
    """unit square with a disk obstacle; exposes the two batch calls the planners need"""
    def __init__(self):
        self.bound = [(0.0, 1.0), (0.0, 1.0)]
        self.eps = 0.01
        self.n_feas = self.n_vis = 0

    def feasible_batch(self, Q):
        Q = np.atleast_2d(Q)
        self.n_feas += 1
        return (((Q - 0.5) ** 2).sum(axis=1) > 0.3 ** 2).astype(np.uint8)

    def visible_batch(self, A, B):
        self.n_vis += 1
        out = np.ones(len(A), dtype=np.uint8)
        for u in np.linspace(0, 1, 101)[1:-1]:
            out &= self.feasible_batch(A * (1 - u) + B * u)
        return out


@pytest.mark.parametrize("kind", ["prm", "lazyprm*"])
def test_roadmap_planner_finds_a_valid_path(kind):
    sp = DiskSpace()
    MotionPlan.setOptions(knn=8, batch=200, seed=3)
    plan = MotionPlan(sp, kind)
    plan.setEndpoints([0.05, 0.5], [0.95, 0.5])
    path = None
    for _ in range(10):
        plan.planMore(1)
        path = plan.getPath()
        if path:
            break
    assert path is not None and path[0] == [0.05, 0.5] and path[-1] == [0.95, 0.5]
    P = np.array(path)
    assert sp.visible_batch(P[:-1], P[1:]).all()                  # every edge of the answer is collision free
    assert plan.pathCost(path) > 0.9 + 0.05                       # must go round the disk
    st = plan.getStats()
    assert st["milestones"] >= 2 and st["samples"] == 200 * st["iterations"]
    V, E = plan.getRoadmap()
    assert len(V) == st["milestones"] and len(E) == st["edges"]
    if kind.startswith("lazy"):
        assert st["edges_checked"] < st["edges"]                  # only edges on candidate paths were validated
    plan.close()


def test_endpoints_must_be_feasible_and_type_checked():
    sp = DiskSpace()
    plan = MotionPlan(sp, "prm")
    with pytest.raises(RuntimeError, match="Start"):
        plan.setEndpoints([0.5, 0.5], [0.9, 0.9])
    with pytest.raises(RuntimeError, match="Goal"):
        plan.setEndpoints([0.1, 0.1], [0.5, 0.45])
    with pytest.raises(ValueError):
        MotionPlan(sp, "fmm")
    with pytest.raises(TypeError):
        MotionPlan(object(), "prm")


@pytest.mark.parametrize("kind", ["rrt", "sbl"])
def test_tree_planners_find_a_valid_path(kind):
    """batched bidirectional RRT / SBL (the planners the north star names next to Lazy-PRM*) on the disk world"""
    sp = DiskSpace()
    MotionPlan.setOptions(batch=64, seed=5, perturbationRadius=0.12)
    plan = MotionPlan(sp, kind)
    plan.setEndpoints([0.05, 0.5], [0.95, 0.5])
    path = None
    for _ in range(60):
        plan.planMore(1)
        path = plan.getPath()
        if path:
            break
    assert path is not None and path[0] == [0.05, 0.5] and path[-1] == [0.95, 0.5]
    P = np.array(path)
    assert sp.visible_batch(P[:-1], P[1:]).all()
    assert plan.pathCost(path) > 0.95
    st = plan.getStats()
    assert st["samples"] == 64 * st["iterations"] and st["milestones"] > 2
    if kind == "sbl":
        assert st["edges_checked"] < st["edges"]                  # tree edges off the answer were never validated
    else:
        V, E = plan.getRoadmap()
        A, B = np.array(V)[[i for i, _ in E]], np.array(V)[[j for _, j in E]]
        assert sp.visible_batch(A, B).all()                       # an RRT only keeps validated edges
    plan.close()


@pytest.mark.parametrize("kind", ["rrt", "lazyprm*"])
def test_shortcutting_after_the_first_plan(kind):
    """MotionPlan.setOptions(shortcut=1) (reference plan/cspace.py): once a path exists, further iterations shorten it with batches of
    chords between points ON the path; the result stays collision free and approaches the taut path round the disk"""
    sp = DiskSpace()
    MotionPlan.setOptions(batch=64, seed=11, perturbationRadius=0.12, knn=8, shortcut=1)
    plan = MotionPlan(sp, kind)
    plan.setEndpoints([0.05, 0.5], [0.95, 0.5])
    for _ in range(80):
        plan.planMore(1)
        if plan.getPath():
            break
    first = plan.getPath()
    assert first is not None
    c0 = plan.pathCost(first)
    costs = [c0]
    for _ in range(15):
        plan.planMore(1)
        costs.append(plan.pathCost(plan.getPath()))
    assert all(b <= a + 1e-12 for a, b in zip(costs[:-1], costs[1:]))      # never longer
    path = plan.getPath()
    P = np.array(path)
    assert path[0] == [0.05, 0.5] and path[-1] == [0.95, 0.5]
    assert sp.visible_batch(P[:-1], P[1:]).all()
    # taut path: two tangents from the endpoints (0.45 from the centre) to the disk of radius 0.3 plus the arc between them
    d, r = 0.45, 0.3
    taut = 2 * np.sqrt(d * d - r * r) + r * (np.pi - 2 * np.arccos(r / d))
    assert taut - 1e-9 <= costs[-1] < min(c0, 1.05 * taut)
    st = plan.getStats()
    assert st["shortcuts_applied"] > 0 and st["shortcuts_tried"] >= st["shortcuts_applied"] and abs(st["path_cost"] - costs[-1]) < 1e-12
    plan.close()


def test_prm_star_neighbourhood_grows_and_cost_converges():
    sp = DiskSpace()
    MotionPlan.setOptions(batch=150, seed=2)
    plan = MotionPlan(sp, "prm*")
    plan.setEndpoints([0.05, 0.5], [0.95, 0.5])
    costs = []
    for _ in range(6):
        plan.planMore(1)
        p = plan.getPath()
        if p:
            costs.append(plan.pathCost(p))
    d, r = 0.45, 0.3
    taut = 2 * np.sqrt(d * d - r * r) + r * (np.pi - 2 * np.arccos(r / d))
    assert len(costs) >= 4 and all(b <= a + 1e-12 for a, b in zip(costs[:-1], costs[1:]))       # a roadmap only gains edges
    assert taut <= costs[-1] < 1.04 * taut
    st = plan.getStats()
    n = st["milestones"]
    assert st["edges_checked"] > n * 4                      # k ~ e (1 + 1/2) log n > 10 neighbours were proposed per vertex
    plan.close()


def test_sensor_ray_construction_follows_the_reference():
    """CameraSensor::GetViewport / the ray-cast fallback (VisualSensors.cpp:424-449,865-897) and LaserRangeSensor's sweep (:57-126):
    rays only, no GPU"""
    import math
    from klampt_b200 import sensing
    cam = sensing.CameraSensor(xres=64, yres=48, xfov=math.radians(90), yfov=math.radians(60), zmin=0.5, zmax=4.0)
    fx, fy, cx, cy = cam.viewport()
    assert fx == pytest.approx(32.0) and fy == pytest.approx(24.0 / math.tan(math.radians(30))) and (cx, cy) == (32.0, 24.0)
    rays, eye, fwd = cam.rays()
    assert rays.shape == (64 * 48, 6) and np.allclose(eye, 0) and np.allclose(fwd, [0, 0, 1])
    k = 24 * 64 + 32                                  # pixel (i, j) = (cx, cy) looks straight ahead and starts zmin along it
    assert np.allclose(rays[k], [0, 0, 0.5, 0, 0, 1])
    k = 24 * 64 + 63                                  # right edge: +x; image rows grow downwards: row 0 looks up = -y of the camera frame
    assert rays[k, 3] > 0.69 and abs(rays[k, 4]) < 1e-12
    assert rays[0, 4] < 0 and rays[47 * 64, 4] > 0
    assert np.allclose(np.linalg.norm(rays[:, 3:], axis=1), 1.0)
    d0 = np.array([0, 0, 1.0]) + (0 - cx) * np.array([1.0, 0, 0]) / fx + (cy - 0) * np.array([0, -1.0, 0]) / fy
    assert np.allclose(rays[0, :3], d0 * 0.5) and np.allclose(rays[0, 3:], d0 / np.linalg.norm(d0))
    las = sensing.LaserRangeSensor(measurementCount=181, depthMinimum=0.1)
    r = las.rays()
    xt, yt = las.angles()
    assert xt[0] == pytest.approx(-math.pi / 2) and xt[90] == pytest.approx(0.0, abs=1e-12) and np.all(yt == 0)
    assert xt[-1] == pytest.approx(-math.pi / 2)      # the sawtooth wraps at u = 1 when the sensor has not been advanced (reference behaviour)
    assert np.allclose(r[90], [0, 0, 0.1, 0, 0, 1])
    las.advance(0.1)
    xt2, _ = las.angles()
    assert xt2[-1] == pytest.approx(math.pi / 2, abs=0.02) and np.all(np.diff(xt2) > 0)
