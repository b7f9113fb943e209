"""Seeded synthetic fixtures for the five workloads of BASELINE.json / SURVEY.md section 8(d).

The data files named by the reference's configs (tx90l.rob, Baxter) live in the separate
Klampt-examples repository, not in the reference tree, so BASELINE.json allows a "procedural 6-link
arm w/ meshes".  Everything here is generated deterministically from ``seed = 20261017 + index`` with
numpy's PCG64 so that the CPU oracle and the GPU engine see bit-identical inputs.

  C1  arm6 + ground + 10 boxes, 10k configs
  C2  arm6 + 200 blob obstacles (~500k triangles), 1M configs
  C3  15-DOF dual-arm torso, self-collision only, 10M configs
  C4  C2 world, 1M straight-line edges at eps = 0.01
  C5  arm6 vs a 5M-point cloud sampled on the C2 obstacle surfaces, collide + distance
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

from .worldspec import (GeomSpec, RobotSpec, WorldSpec, REVOLUTE, JOINT_NORMAL, JOINT_WELD, JOINT_FLOATING, JOINT_BALLANDSOCKET, IDENTITY12)

BASE_SEED = 20261017


# --------------------------------------------------------------------------------------- meshes
def unit_cube() -> Tuple[np.ndarray, np.ndarray]:
    """The unit cube [0,1]^3 as 8 vertices / 12 triangles (same solid as the reference's only mesh
    asset, tests/objects/cube.off)."""
    v = np.array([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)], dtype=np.float64)
    t = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                  [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    return v, t


def box_mesh(lo, hi, div=1) -> Tuple[np.ndarray, np.ndarray]:
    """Axis-aligned box with each face split into div x div quads (2 triangles each)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    verts, tris = [], []
    for ax in range(3):
        u, w = (ax + 1) % 3, (ax + 2) % 3
        for side in (0, 1):
            base = len(verts)
            for i in range(div + 1):
                for j in range(div + 1):
                    p = np.zeros(3)
                    p[ax] = hi[ax] if side else lo[ax]
                    p[u] = lo[u] + (hi[u] - lo[u]) * i / div
                    p[w] = lo[w] + (hi[w] - lo[w]) * j / div
                    verts.append(p)
            for i in range(div):
                for j in range(div):
                    a = base + i * (div + 1) + j
                    b = a + 1
                    c = a + (div + 1)
                    d = c + 1
                    if side:
                        tris += [[a, c, d], [a, d, b]]
                    else:
                        tris += [[a, d, c], [a, b, d]]
    return np.array(verts, dtype=np.float64), np.array(tris, dtype=np.int32)


def icosphere(subdiv: int) -> Tuple[np.ndarray, np.ndarray]:
    """Unit icosphere, 20*4^subdiv triangles."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    verts = [np.array(p, dtype=np.float64) / math.sqrt(1 + t * t) for p in v]
    faces = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
             [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
             [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        faces = nf
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int32)


_ICO_CACHE = {}


def blob_mesh(rng: np.random.Generator, subdiv: int, radius: float, noise: float = 0.25) -> Tuple[np.ndarray, np.ndarray]:
    """Convex-ish blob: icosphere with smooth low-frequency radial noise."""
    if subdiv not in _ICO_CACHE:
        _ICO_CACHE[subdiv] = icosphere(subdiv)
    v, t = _ICO_CACHE[subdiv]
    dirs = rng.normal(size=(4, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    amp = rng.uniform(-noise, noise, size=4) / 2.0
    freq = rng.integers(1, 4, size=4)
    r = np.ones(len(v))
    for d, a, f in zip(dirs, amp, freq):
        r += a * np.cos(f * np.arccos(np.clip(v @ d, -1, 1)))
    return v * (radius * r)[:, None], t.copy()


def capsule_mesh(r: float, z0: float, z1: float, nseg: int = 24, ncap: int = 6, nbody: int = 8) -> Tuple[np.ndarray, np.ndarray]:
    """Capsule of radius r around the local z axis from z0 to z1 (hemispherical caps included)."""
    rings = []
    for i in range(1, ncap + 1):                                    # bottom cap, pole excluded
        a = -math.pi / 2 + (math.pi / 2) * i / ncap
        rings.append((r * math.cos(a), z0 + r * math.sin(a)))
    for i in range(1, nbody + 1):
        rings.append((r, z0 + (z1 - z0) * i / nbody))
    for i in range(1, ncap):
        a = (math.pi / 2) * i / ncap
        rings.append((r * math.cos(a), z1 + r * math.sin(a)))
    verts = [[0.0, 0.0, z0 - r]]
    for rr, z in rings:
        for s in range(nseg):
            th = 2 * math.pi * s / nseg
            verts.append([rr * math.cos(th), rr * math.sin(th), z])
    verts.append([0.0, 0.0, z1 + r])
    top = len(verts) - 1
    tris = []
    for s in range(nseg):
        tris.append([0, 1 + (s + 1) % nseg, 1 + s])
    for k in range(len(rings) - 1):
        b0, b1 = 1 + k * nseg, 1 + (k + 1) * nseg
        for s in range(nseg):
            s1 = (s + 1) % nseg
            tris += [[b0 + s, b0 + s1, b1 + s1], [b0 + s, b1 + s1, b1 + s]]
    bl = 1 + (len(rings) - 1) * nseg
    for s in range(nseg):
        tris.append([bl + s, bl + (s + 1) % nseg, top])
    return np.array(verts, dtype=np.float64), np.array(tris, dtype=np.int32)


def transform_points(T12: np.ndarray, p: np.ndarray) -> np.ndarray:
    R = np.asarray(T12[:9]).reshape(3, 3)
    return p @ R.T + np.asarray(T12[9:12])


def rot_axis_angle(axis, angle) -> np.ndarray:
    w = np.asarray(axis, dtype=np.float64)
    w = w / np.linalg.norm(w)
    c, s = math.cos(angle), math.sin(angle)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return c * np.eye(3) + (1 - c) * np.outer(w, w) + s * K


def make_T(R=None, t=(0, 0, 0)) -> np.ndarray:
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    return np.concatenate([R.reshape(-1), np.asarray(t, dtype=np.float64)])


def merge_meshes(parts):
    vs, ts, off = [], [], 0
    for v, t in parts:
        vs.append(v)
        ts.append(t + off)
        off += len(v)
    return np.vstack(vs), np.vstack(ts).astype(np.int32)


def rotate_mesh(v: np.ndarray, R: np.ndarray, t=(0, 0, 0)) -> np.ndarray:
    return v @ np.asarray(R).T + np.asarray(t, dtype=np.float64)


# --------------------------------------------------------------------------------------- robots
def make_arm6(world: WorldSpec) -> RobotSpec:
    """Procedural 6-DOF arm, TX90L-like proportions (reach ~1.2 m): a welded base link + 6 revolute
    links, axes z,y,y,z,y,z, parents [-1,0,1,2,3,4,5], capsule/box link meshes (~8k triangles)."""
    Ry90 = rot_axis_angle([0, 1, 0], math.pi / 2)
    geoms = []
    # link 0: pedestal (welded); box 0.28 x 0.28 x 0.22 sitting on the ground
    geoms.append(box_mesh([-0.14, -0.14, 0.0], [0.14, 0.14, 0.22], div=8))
    # link 1: shoulder column (yaw about z)
    geoms.append(capsule_mesh(0.105, 0.02, 0.16, nseg=28, ncap=7, nbody=6))
    # link 2: upper arm 0.5 m along local z, offset sideways in y
    v, t = capsule_mesh(0.075, 0.0, 0.50, nseg=28, ncap=7, nbody=16)
    geoms.append((v + np.array([0.0, 0.15, 0.0]), t))
    # link 3: elbow block
    geoms.append(capsule_mesh(0.07, -0.02, 0.10, nseg=24, ncap=6, nbody=4))
    # link 4: forearm 0.40 m along local z
    geoms.append(capsule_mesh(0.055, 0.16, 0.47, nseg=24, ncap=6, nbody=12))
    # link 5: wrist (pitch about y), short capsule across y
    v, t = capsule_mesh(0.045, -0.03, 0.03, nseg=20, ncap=5, nbody=2)
    geoms.append((rotate_mesh(v, rot_axis_angle([1, 0, 0], math.pi / 2)), t))
    # link 6: flange + tool stub along z
    v1, t1 = capsule_mesh(0.032, 0.075, 0.16, nseg=16, ncap=4, nbody=4)
    v2, t2 = box_mesh([-0.06, -0.02, 0.17], [0.06, 0.02, 0.23], div=3)
    geoms.append(merge_meshes([(v1, t1), (v2, t2)]))
    del Ry90
    link_geom = [world.add_geom(GeomSpec.mesh(v, t)) for v, t in geoms]
    T0 = np.tile(IDENTITY12, (7, 1))
    T0[1, 9:12] = [0.0, 0.0, 0.22]      # shoulder yaw sits on the pedestal
    T0[2, 9:12] = [0.05, 0.0, 0.26]     # shoulder pitch
    T0[3, 9:12] = [0.0, 0.15, 0.50]     # elbow pitch at the end of the upper arm
    T0[4, 9:12] = [0.0, -0.10, 0.0]     # forearm roll, back towards the arm plane
    T0[5, 9:12] = [0.0, 0.0, 0.55]      # wrist pitch
    T0[6, 9:12] = [0.0, 0.0, 0.0]       # flange roll
    axis = np.array([[0, 0, 1], [0, 0, 1], [0, 1, 0], [0, 1, 0], [0, 0, 1], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    qmin = np.array([0.0, -math.pi, -2.27, -2.5, -math.pi, -2.1, -math.pi])
    qmax = np.array([0.0, math.pi, 2.27, 2.5, math.pi, 2.3, math.pi])
    return RobotSpec(parents=np.array([-1, 0, 1, 2, 3, 4, 5], dtype=np.int32), linktype=np.full(7, REVOLUTE, dtype=np.uint8),
                     axis=axis, T0=T0, qmin=qmin, qmax=qmax, link_geom=link_geom,
                     joint_type=np.array([JOINT_WELD] + [JOINT_NORMAL] * 6, dtype=np.uint8),
                     joint_link=np.arange(7, dtype=np.int32),
                     names=["base", "shoulder", "upperarm", "elbow", "forearm", "wrist", "flange"])


def make_dualarm15(world: WorldSpec) -> RobotSpec:
    """15-DOF dual-arm torso: welded pedestal, torso pan, then two 7-DOF arms whose shoulders sit 0.5 m
    apart, each preceded by a welded mounting frame; 19 links, 15 moving DOF."""
    parents, axes, T0s, qmins, qmaxs, jtypes, geoms, names = [], [], [], [], [], [], [], []

    def add(parent, axis, t, R, lo, hi, jt, mesh, name):
        parents.append(parent)
        axes.append(axis)
        T0s.append(make_T(R, t))
        qmins.append(lo)
        qmaxs.append(hi)
        jtypes.append(jt)
        geoms.append(mesh)
        names.append(name)
        return len(parents) - 1

    ped = add(-1, [0, 0, 1], [0, 0, 0], None, 0.0, 0.0, JOINT_WELD, box_mesh([-0.2, -0.2, 0.0], [0.2, 0.2, 0.9], div=8), "pedestal")
    torso = add(ped, [0, 0, 1], [0, 0, 0.9], None, -1.6, 1.6, JOINT_NORMAL,
                capsule_mesh(0.16, 0.05, 0.35, nseg=32, ncap=8, nbody=8), "torso")
    for side, sgn in (("left", 1.0), ("right", -1.0)):
        Rm = rot_axis_angle([1, 0, 0], -sgn * math.pi / 4)          # arms mounted tilted outwards
        mount = add(torso, [0, 0, 1], [0.06, sgn * 0.25, 0.30], Rm, 0.0, 0.0, JOINT_WELD, None, side + "_mount")
        s0 = add(mount, [0, 0, 1], [0, 0, 0.0], None, -1.7, 1.7, JOINT_NORMAL,
                 capsule_mesh(0.07, 0.0, 0.12, nseg=24, ncap=6, nbody=4), side + "_s0")
        s1 = add(s0, [0, 1, 0], [0.07, 0, 0.17], None, -2.1, 1.0, JOINT_NORMAL,
                 capsule_mesh(0.065, 0.0, 0.06, nseg=24, ncap=6, nbody=3), side + "_s1")
        e0 = add(s1, [0, 0, 1], [0, 0, 0.10], None, -3.0, 3.0, JOINT_NORMAL,
                 capsule_mesh(0.06, 0.02, 0.26, nseg=24, ncap=6, nbody=10), side + "_e0")
        e1 = add(e0, [0, 1, 0], [0.07, 0, 0.36], None, -0.05, 2.6, JOINT_NORMAL,
                 capsule_mesh(0.055, 0.0, 0.05, nseg=24, ncap=6, nbody=3), side + "_e1")
        w0 = add(e1, [0, 0, 1], [0, 0, 0.10], None, -3.0, 3.0, JOINT_NORMAL,
                 capsule_mesh(0.05, 0.02, 0.27, nseg=24, ncap=6, nbody=10), side + "_w0")
        w1 = add(w0, [0, 1, 0], [0.01, 0, 0.37], None, -1.57, 2.0, JOINT_NORMAL,
                 capsule_mesh(0.042, 0.0, 0.04, nseg=20, ncap=5, nbody=2), side + "_w1")
        add(w1, [0, 0, 1], [0, 0, 0.09], None, -3.0, 3.0, JOINT_NORMAL,
            merge_meshes([capsule_mesh(0.035, 0.0, 0.06, nseg=16, ncap=4, nbody=3),
                          box_mesh([-0.05, -0.015, 0.10], [0.05, 0.015, 0.19], div=3)]), side + "_w2")
    link_geom = [(-1 if m is None else world.add_geom(GeomSpec.mesh(*m))) for m in geoms]
    L = len(parents)
    return RobotSpec(parents=np.array(parents, dtype=np.int32), linktype=np.full(L, REVOLUTE, dtype=np.uint8),
                     axis=np.array(axes, dtype=np.float64), T0=np.array(T0s), qmin=np.array(qmins), qmax=np.array(qmaxs),
                     link_geom=link_geom, joint_type=np.array(jtypes, dtype=np.uint8), joint_link=np.arange(L, dtype=np.int32),
                     names=names)


def make_planar_nR(world: WorldSpec, n: int, link_length: float = 1.0) -> RobotSpec:
    """Planar nR arm in the style of the reference's procedural template
    (Python/klampt/model/create/planar_robot.py:20-70): every link rotates about y, offset link_length
    along x from its parent, box geometry 1 x 0.1 x 0.1 scaled by the link length.  Closed-form FK."""
    v, t = box_mesh([0.0, -0.05, -0.05], [1.0, 0.05, 0.05], div=1)
    v = v * max(link_length, 0.05)
    T0 = np.tile(IDENTITY12, (n, 1))
    T0[1:, 9] = link_length
    link_geom = [world.add_geom(GeomSpec.mesh(v, t)) for _ in range(n)]
    return RobotSpec(parents=np.arange(-1, n - 1, dtype=np.int32), linktype=np.full(n, REVOLUTE, dtype=np.uint8),
                     axis=np.tile(np.array([0.0, 1.0, 0.0]), (n, 1)), T0=T0, qmin=np.zeros(n), qmax=np.full(n, 6.28319),
                     link_geom=link_geom, joint_type=np.full(n, JOINT_NORMAL, dtype=np.uint8),
                     joint_link=np.arange(n, dtype=np.int32))


def make_floating_body(world: WorldSpec, with_ball_wrist: bool = True) -> RobotSpec:
    """Free-flying body the way Klamp't models one (Cpp/docs/Manual-FileTypes.md "joint floating"): three prismatic
    virtual links (x, y, z) and three revolute ones (z, y, x) with no geometry of their own, the body mesh on the
    last; then a one-link arm on a Normal joint and, optionally, a tool on a ball-and-socket joint (three revolute
    links about z, y, x).  Joints: Floating(link 5, base -1), Normal(6), BallAndSocket(link 9, base 6)."""
    P, R = 1, 0
    parents, ltype, axes, T0s, qmin, qmax, geoms = [], [], [], [], [], [], []
    def add(parent, lt, axis, t, lo, hi, mesh):
        parents.append(parent); ltype.append(lt); axes.append(axis); T0s.append(make_T(None, t)); qmin.append(lo); qmax.append(hi)
        geoms.append(-1 if mesh is None else world.add_geom(GeomSpec.mesh(*mesh)))
        return len(parents) - 1
    l = add(-1, P, [1, 0, 0], (0, 0, 1.0), -1.0, 1.0, None)
    l = add(l, P, [0, 1, 0], (0, 0, 0), -1.0, 1.0, None)
    l = add(l, P, [0, 0, 1], (0, 0, 0), -0.6, 0.8, None)
    l = add(l, R, [0, 0, 1], (0, 0, 0), -math.pi, math.pi, None)
    l = add(l, R, [0, 1, 0], (0, 0, 0), -1.5, 1.5, None)
    body = add(l, R, [1, 0, 0], (0, 0, 0), -math.pi, math.pi, box_mesh([-0.25, -0.15, -0.1], [0.25, 0.15, 0.1], div=6))
    arm = add(body, R, [0, 1, 0], (0.25, 0, 0), -2.0, 2.0, capsule_mesh(0.05, 0.08, 0.40, nseg=16, ncap=4, nbody=6))
    jt, jl, jb = [JOINT_FLOATING, JOINT_NORMAL], [body, arm], [-1, body]
    if with_ball_wrist:
        l = add(arm, R, [0, 0, 1], (0, 0, 0.47), -math.pi, math.pi, None)
        l = add(l, R, [0, 1, 0], (0, 0, 0), -1.5, 1.5, None)
        tool = add(l, R, [1, 0, 0], (0, 0, 0), -math.pi, math.pi, box_mesh([-0.03, -0.08, 0.02], [0.03, 0.08, 0.2], div=3))
        jt.append(JOINT_BALLANDSOCKET); jl.append(tool); jb.append(arm)
    return RobotSpec(parents=np.array(parents, dtype=np.int32), linktype=np.array(ltype, dtype=np.uint8),
                     axis=np.array(axes, dtype=np.float64), T0=np.array(T0s), qmin=np.array(qmin), qmax=np.array(qmax),
                     link_geom=geoms, joint_type=np.array(jt, dtype=np.uint8), joint_link=np.array(jl, dtype=np.int32),
                     joint_base=np.array(jb, dtype=np.int32))


# --------------------------------------------------------------------------------------- worlds
def _ground(world: WorldSpec, half=2.0, div=8):
    v, t = box_mesh([-half, -half, -0.05], [half, half, 0.0], div=div)
    world.terrains.append(world.add_geom(GeomSpec.mesh(v, t)))


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _obstacle_centres(rng, n, keepout=0.25, box=((-1.5, 1.5), (-1.5, 1.5), (0.0, 2.0))):
    out = []
    while len(out) < n:
        c = np.array([rng.uniform(*box[0]), rng.uniform(*box[1]), rng.uniform(*box[2])])
        if math.hypot(c[0], c[1]) < keepout:
            continue
        out.append(c)
    return out


def world_c1(seed_index: int = 1) -> WorldSpec:
    """C1: arm6 + ground slab + 10 random boxes."""
    rng = np.random.default_rng(BASE_SEED + seed_index)
    w = WorldSpec()
    _ground(w)
    for c in _obstacle_centres(rng, 10, keepout=0.45):
        d = rng.uniform(0.08, 0.25, size=3)
        v, t = box_mesh(-d, d, div=4)
        w.objects.append((w.add_geom(GeomSpec.mesh(v, t)), make_T(_random_rotation(rng), c)))
    w.robot = make_arm6(w)
    return w


def world_c2(seed_index: int = 2, n_obstacles: int = 200, fine_fraction: float = 0.32, keepout: float = 0.40,
             scale=(0.04, 0.16)) -> WorldSpec:
    """C2: arm6 in a cluttered world: n_obstacles blobs (icosphere subdiv 3 or 4 with radial noise; 200 of them
    give ~500k triangles), centres uniform in a 3 x 3 x 2 m box minus a keep-out cylinder round the base,
    radius U(scale) m, each a rigid object with a random pose; plus the ground slab as a terrain.
    (SURVEY.md 8d proposed radius U(0.05,0.3) / keep-out 0.25 m; with this arm that leaves only 9 % of the
    configurations feasible, outside the 30-70 % infeasible range the survey aims for, so the defaults here
    are radius U(0.04,0.16) / keep-out 0.40 m, which gives ~57 % infeasible.  DESIGN.md records this.)"""
    rng = np.random.default_rng(BASE_SEED + seed_index)
    w = WorldSpec()
    _ground(w)
    for c in _obstacle_centres(rng, n_obstacles, keepout=keepout):
        sub = 4 if rng.uniform() < fine_fraction else 3
        # the blob must not swallow the robot base: shrink those that reach into the keep-out cylinder
        r = rng.uniform(*scale)
        r = min(r, max(0.03, (math.hypot(c[0], c[1]) - 0.16) / 1.3))
        v, t = blob_mesh(rng, sub, r)
        w.objects.append((w.add_geom(GeomSpec.mesh(v, t)), make_T(_random_rotation(rng), c)))
    w.robot = make_arm6(w)
    return w


def world_floating(seed_index: int = 7, n_obstacles: int = 40) -> WorldSpec:
    """Free-flying body + arm + ball-and-socket tool among blobs over the ground slab (exercises Floating /
    BallAndSocket joints in the edge metric and interpolation)."""
    rng = np.random.default_rng(BASE_SEED + seed_index)
    w = WorldSpec()
    _ground(w)
    for c in _obstacle_centres(rng, n_obstacles, keepout=0.0):
        v, t = blob_mesh(rng, 3, rng.uniform(0.05, 0.2))
        w.objects.append((w.add_geom(GeomSpec.mesh(v, t)), make_T(_random_rotation(rng), c)))
    w.robot = make_floating_body(w)
    return w


def world_boxes(seed_index: int = 8, n_boxes: int = 14, n_blobs: int = 6) -> WorldSpec:
    """arm6 whose pedestal and elbow are solid box primitives, among solid oriented boxes (some large enough to swallow a
    whole link: only the solid semantics sees those), a few blobs, over a solid slab (AABB primitive) as the terrain."""
    rng = np.random.default_rng(BASE_SEED + seed_index)
    w = WorldSpec()
    w.terrains.append(w.add_geom(GeomSpec.aabb([-2.0, -2.0, -0.05], [2.0, 2.0, 0.0])))
    for i, c in enumerate(_obstacle_centres(rng, n_boxes, keepout=0.35)):
        big = i % 4 == 0
        h = rng.uniform(0.15, 0.3, size=3) if big else rng.uniform(0.03, 0.12, size=3)
        g = w.add_geom(GeomSpec.box(rng.uniform(-0.02, 0.02, size=3), _random_rotation(rng), h))
        w.objects.append((g, make_T(_random_rotation(rng), c)))
    for c in _obstacle_centres(rng, n_blobs, keepout=0.4):
        v, t = blob_mesh(rng, 3, rng.uniform(0.05, 0.15))
        w.objects.append((w.add_geom(GeomSpec.mesh(v, t)), make_T(_random_rotation(rng), c)))
    w.robot = make_arm6(w)
    w.robot.link_geom[0] = w.add_geom(GeomSpec.aabb([-0.14, -0.14, 0.0], [0.14, 0.14, 0.22]))
    w.robot.link_geom[3] = w.add_geom(GeomSpec.box([0.0, 0.0, 0.04], rot_axis_angle([0, 0, 1], 0.3), [0.07, 0.06, 0.08]))
    return w


def world_c3() -> WorldSpec:
    """C3: 15-DOF dual-arm torso, no environment (self-collision only)."""
    w = WorldSpec()
    w.robot = make_dualarm15(w)
    return w


def world_c5(seed_index: int = 5, n_points: int = 5_000_000, n_obstacles: int = 200, noise: float = 0.005,
             margin: float = 0.005) -> WorldSpec:
    """C5: arm6 meshes vs one point cloud sampled on the C2 obstacle surfaces + 5 mm Gaussian noise
    (point radius 0, collision margin 5 mm), as a single terrain."""
    src = world_c2(2, n_obstacles)
    rng = np.random.default_rng(BASE_SEED + seed_index)
    areas, tris_w = [], []
    for gi, T in src.objects:
        g = src.geoms[gi]
        v = transform_points(T, g.verts)
        tw = v[g.tris]                                      # (nt,3,3)
        tris_w.append(tw)
    tw = np.concatenate(tris_w)
    areas = 0.5 * np.linalg.norm(np.cross(tw[:, 1] - tw[:, 0], tw[:, 2] - tw[:, 0]), axis=1)
    pick = rng.choice(len(tw), size=n_points, p=areas / areas.sum())
    u = rng.uniform(size=(n_points, 2))
    su = np.sqrt(u[:, 0])
    b0, b1, b2 = 1 - su, su * (1 - u[:, 1]), su * u[:, 1]
    pts = tw[pick, 0] * b0[:, None] + tw[pick, 1] * b1[:, None] + tw[pick, 2] * b2[:, None]
    pts += rng.normal(scale=noise, size=pts.shape)
    w = WorldSpec()
    w.terrains.append(w.add_geom(GeomSpec.cloud(pts, None, margin=margin)))
    w.robot = make_arm6(w)
    return w


# --------------------------------------------------------------------------------------- samplers
def sample_configs(robot: RobotSpec, n: int, seed_index: int) -> np.ndarray:
    """Uniform per joint in [qmin,qmax] (RobotCSpace::Sample for Normal joints,
    Cpp/Planning/RobotCSpace.cpp:85-87)."""
    rng = np.random.default_rng(BASE_SEED + 100 + seed_index)
    return rng.uniform(robot.qmin, robot.qmax, size=(n, robot.L))


def sample_edges(robot: RobotSpec, feasible_fn, n: int, seed_index: int, rmin=0.2, rmax=2.0):
    """C4 edges: a = uniform *feasible* config (rejection-sampled with ``feasible_fn(Q)->bool array``),
    b = a + delta, delta uniform in a ball of C-space radius U(rmin,rmax), clamped to the limits."""
    rng = np.random.default_rng(BASE_SEED + 200 + seed_index)
    A = np.empty((0, robot.L))
    while len(A) < n:
        Q = rng.uniform(robot.qmin, robot.qmax, size=(max(1024, 2 * (n - len(A))), robot.L))
        ok = np.asarray(feasible_fn(Q)).astype(bool)
        A = np.vstack([A, Q[ok]])
    A = A[:n]
    moving = robot.qmax > robot.qmin
    d = rng.normal(size=(n, robot.L)) * moving
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = rng.uniform(rmin, rmax, size=n) * rng.uniform(size=n) ** (1.0 / max(1, int(moving.sum())))
    B = np.clip(A + d * rad[:, None], robot.qmin, robot.qmax)
    return np.ascontiguousarray(A), np.ascontiguousarray(B)
