"""ctypes front end of the CPU oracle (oracle/kb_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from klampt_b200/.  PARITY UNPINNED -- see kb_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


class Counts(C.Structure):
    _fields_ = [("n_box", C.c_int64), ("n_node", C.c_int64), ("n_tri", C.c_int64), ("n_pt", C.c_int64)]


COUNTS_DTYPE = np.dtype([("n_box", np.int64), ("n_node", np.int64), ("n_tri", np.int64), ("n_pt", np.int64)])


def build(force: bool = False, variant: str = "strict") -> str:
    """strict: the checker (median-split tree, no FMA contraction, portable ISA).  fast: the CPU-baseline arm of bench.py (binned
    SAH tree, -march=native, FMA allowed) -- always rebuilt on the machine that runs it, because -march=native does not travel."""
    fast = variant == "fast"
    so = os.path.join(_HERE, "libkb_oracle_fast.so" if fast else "libkb_oracle.so")
    src = os.path.join(_HERE, "kb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["cleanfast", "fast"] if fast else ["clean", "all"]))
    return so


def lib(variant: str = "strict"):
    if variant in _LIBS:
        return _LIBS[variant]
    so = build(force=(variant == "fast"), variant=variant)
    try:
        L = C.CDLL(so)
    except OSError:
        so = build(force=True, variant=variant)
        L = C.CDLL(so)
    dp, ip, u8p, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
    L.ko_create.restype = vp
    L.ko_destroy.argtypes = [vp]
    L.ko_add_trimesh.argtypes = [vp, dp, C.c_int, ip, C.c_int, C.c_double]
    L.ko_add_pointcloud.argtypes = [vp, dp, C.c_int, dp, C.c_double]
    L.ko_add_primitive.argtypes = [vp, C.c_int, dp, C.c_double]
    L.ko_add_terrain.argtypes = [vp, C.c_int]
    L.ko_add_rigid_object.argtypes = [vp, C.c_int, dp]
    L.ko_robot_create.argtypes = [vp, C.c_int, ip, u8p, dp, dp, dp, dp]
    L.ko_robot_set_link_geometry.argtypes = [vp, C.c_int, C.c_int]
    L.ko_robot_set_joints.argtypes = [vp, C.c_int, u8p, ip, ip]
    L.ko_robot_add_affine_driver.argtypes = [vp, C.c_int, ip, dp, dp, C.c_double, C.c_double]
    L.ko_robot_set_self_collision.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.ko_set_pair_mask.argtypes = [vp, u8p, C.c_int]
    L.ko_finalize.argtypes = [vp]
    L.ko_num_ids.argtypes = [vp]
    L.ko_get_pair_mask.argtypes = [vp, u8p]
    L.ko_fk.argtypes = [vp, dp, dp]
    L.ko_check_joint_limits.argtypes = [vp, dp]
    L.ko_feasible.argtypes = [vp, dp, ip, C.POINTER(Counts)]
    L.ko_feasible_brute.argtypes = [vp, dp]
    L.ko_feasible_batch.argtypes = [vp, dp, C.c_int64, u8p, ip, vp, C.c_int]
    L.ko_cspace_distance.argtypes = [vp, dp, dp, dp]
    L.ko_cspace_distance.restype = C.c_double
    L.ko_interpolate.argtypes = [vp, dp, dp, C.c_double, dp]
    L.ko_edge_visible.argtypes = [vp, dp, dp, C.c_double, dp, ip, C.POINTER(Counts)]
    L.ko_edges_visible_batch.argtypes = [vp, dp, dp, C.c_int64, C.c_double, dp, u8p, ip, C.c_int]
    L.ko_distance.argtypes = [vp, dp, C.c_double, C.c_int, ip, C.POINTER(Counts)]
    L.ko_distance.restype = C.c_double
    L.ko_distance_batch.argtypes = [vp, dp, C.c_int64, C.c_double, C.c_int, dp, ip, C.c_int]
    L.ko_geom_collides.argtypes = [vp, C.c_int, dp, C.c_int, dp]
    L.ko_geom_within_distance.argtypes = [vp, C.c_int, dp, C.c_int, dp, C.c_double]
    L.ko_geom_distance.argtypes = [vp, C.c_int, dp, C.c_int, dp, C.c_double]
    L.ko_geom_distance.restype = C.c_double
    L.ko_geom_distance_brute.argtypes = [vp, C.c_int, dp, C.c_int, dp]
    L.ko_geom_distance_brute.restype = C.c_double
    L.ko_geom_aabb.argtypes = [vp, C.c_int, dp, dp, dp]
    L.ko_tri_tri_intersect.argtypes = [dp, dp]
    L.ko_tri_tri_distance.argtypes = [dp, dp]
    L.ko_tri_tri_distance.restype = C.c_double
    L.ko_point_tri_distance.argtypes = [dp, dp]
    L.ko_point_tri_distance.restype = C.c_double
    L.ko_seg_seg_distance.argtypes = [dp, dp, dp, dp]
    L.ko_seg_seg_distance.restype = C.c_double
    L.ko_max_threads.restype = C.c_int
    L.ko_penetration.argtypes = [vp, dp, C.c_int]
    L.ko_penetration.restype = C.c_double
    L.ko_tri_tri_depth.argtypes = [dp, dp]
    L.ko_tri_tri_depth.restype = C.c_double
    L.ko_geom_penetration.argtypes = [vp, C.c_int, dp, C.c_int, dp, C.c_double]
    L.ko_geom_penetration.restype = C.c_double
    L.ko_raycast.argtypes = [vp, dp, dp, dp, u8p, dp, ip]
    L.ko_raycast_batch.argtypes = [vp, dp, dp, C.c_int64, u8p, ip, dp, ip, C.c_int]
    L.ko_geom_raycast.argtypes = [vp, C.c_int, dp, dp, dp, dp, ip, C.c_int]
    L.ko_raycast_counts.argtypes = [vp, dp, dp, C.c_int64, C.POINTER(Counts)]
    _LIBS[variant] = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def tri_tri_intersect(a, b) -> bool:
    a_, ap = _d(np.asarray(a).reshape(9))
    b_, bp = _d(np.asarray(b).reshape(9))
    return bool(lib().ko_tri_tri_intersect(ap, bp))


def tri_tri_distance(a, b) -> float:
    a_, ap = _d(np.asarray(a).reshape(9))
    b_, bp = _d(np.asarray(b).reshape(9))
    return float(lib().ko_tri_tri_distance(ap, bp))


def point_tri_distance(p, t) -> float:
    p_, pp = _d(np.asarray(p).reshape(3))
    t_, tp = _d(np.asarray(t).reshape(9))
    return float(lib().ko_point_tri_distance(pp, tp))


def seg_seg_distance(p0, p1, q0, q1) -> float:
    a = [_d(np.asarray(x).reshape(3)) for x in (p0, p1, q0, q1)]
    return float(lib().ko_seg_seg_distance(*[x[1] for x in a]))


def tri_tri_depth(a, b) -> float:
    a_, ap = _d(np.asarray(a).reshape(9))
    b_, bp = _d(np.asarray(b).reshape(9))
    return float(lib().ko_tri_tri_depth(ap, bp))


def max_threads() -> int:
    return int(lib().ko_max_threads())


class OracleWorld:
    """Builds the oracle's world from a klampt_b200.worldspec.WorldSpec."""

    def __init__(self, spec, variant: str = "strict"):
        L = lib(variant)
        self.L = L
        self.variant = variant
        self.spec = spec
        self.h = L.ko_create()
        for g in spec.geoms:
            if g.kind == "mesh":
                v, vp = _d(g.verts)
                t, tp = _i(g.tris)
                L.ko_add_trimesh(self.h, vp, len(v), tp, len(t), g.margin)
            elif g.kind == "cloud":
                p, pp = _d(g.points)
                if g.radius is None:
                    L.ko_add_pointcloud(self.h, pp, len(p), None, g.margin)
                else:
                    r, rp = _d(g.radius)
                    L.ko_add_pointcloud(self.h, pp, len(p), rp, g.margin)
            elif g.kind in ("triangle", "box", "segment"):
                p, pp = _d(g.params)
                if L.ko_add_primitive(self.h, {"triangle": 2, "box": 3, "segment": 5}[g.kind], pp, g.margin) < 0:
                    raise ValueError("the oracle rejected a %s primitive" % g.kind)
            elif g.kind in ("sphere", "point"):
                p, pp = _d(g.params)
                L.ko_add_primitive(self.h, 1 if g.kind == "sphere" else 0, pp, g.margin)
            else:
                t, tp = _i(np.zeros((0, 3)))
                v, vp = _d(np.zeros((0, 3)))
                L.ko_add_trimesh(self.h, vp, 0, tp, 0, 0.0)
        for gi in spec.terrains:
            L.ko_add_terrain(self.h, gi)
        for gi, T in spec.objects:
            T_, Tp = _d(T)
            L.ko_add_rigid_object(self.h, gi, Tp)
        r = spec.robot
        self.nlinks = 0
        if r is not None:
            self.nlinks = r.L
            a = [_i(r.parents), _u8(r.linktype), _d(r.axis), _d(r.T0), _d(r.qmin), _d(r.qmax)]
            rc = L.ko_robot_create(self.h, r.L, *[x[1] for x in a])
            if rc != 0:
                raise ValueError("ko_robot_create failed: %d" % rc)
            for j, gi in enumerate(r.link_geom):
                L.ko_robot_set_link_geometry(self.h, j, gi)
            if r.joint_type is not None:
                jt, jtp = _u8(r.joint_type)
                jl, jlp = _i(r.joint_link)
                jbp = None
                if getattr(r, "joint_base", None) is not None:
                    jb, jbp = _i(r.joint_base)
                if L.ko_robot_set_joints(self.h, len(jt), jtp, jlp, jbp) != 0:
                    raise ValueError("ko_robot_set_joints: a Floating / FloatingPlanar / BallAndSocket joint does not drive the link chain the reference asserts")
            for d in r.drivers:
                li, lip = _i(d.links)
                sc, scp = _d(d.scale)
                of, ofp = _d(d.offset)
                L.ko_robot_add_affine_driver(self.h, len(li), lip, scp, ofp, d.qmin, d.qmax)
            for (i, j, en) in r.self_collision_edits:
                L.ko_robot_set_self_collision(self.h, i, j, int(en))
        if spec.pair_mask is not None:
            m, mp = _u8(spec.pair_mask)
            if L.ko_set_pair_mask(self.h, mp, m.shape[0]) != 0:
                raise ValueError("pair mask size mismatch")
        L.ko_finalize(self.h)

    def close(self):
        if self.h:
            self.L.ko_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ queries
    def num_ids(self) -> int:
        return int(self.L.ko_num_ids(self.h))

    def pair_mask(self) -> np.ndarray:
        n = self.num_ids()
        m = np.zeros((n, n), dtype=np.uint8)
        self.L.ko_get_pair_mask(self.h, m.ctypes.data_as(C.POINTER(C.c_uint8)))
        return m

    def fk(self, q) -> np.ndarray:
        q_, qp = _d(q)
        T = np.zeros((self.nlinks, 12))
        self.L.ko_fk(self.h, qp, T.ctypes.data_as(C.POINTER(C.c_double)))
        return T

    def fk_batch(self, Q) -> np.ndarray:
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nlinks)
        return np.stack([self.fk(q) for q in Q])

    def check_joint_limits(self, q) -> bool:
        q_, qp = _d(q)
        return bool(self.L.ko_check_joint_limits(self.h, qp))

    def feasible(self, q, want_counts=False):
        q_, qp = _d(q)
        pr = np.zeros(2, dtype=np.int32)
        cnt = Counts()
        ok = bool(self.L.ko_feasible(self.h, qp, pr.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cnt)))
        if want_counts:
            return ok, (int(pr[0]), int(pr[1])), cnt
        return ok

    def penetration(self, q, include_self=True) -> float:
        """deepest contact of q in metres (-1: nothing in contact); the colliding side of the two-sided 1e-6 m band"""
        q_, qp = _d(q)
        return float(self.L.ko_penetration(self.h, qp, int(include_self)))

    def feasible_brute(self, q) -> bool:
        q_, qp = _d(q)
        return bool(self.L.ko_feasible_brute(self.h, qp))

    def feasible_batch(self, Q, nthreads=0, want_pairs=False, want_counts=False):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nlinks)
        N = Q.shape[0]
        out = np.zeros(N, dtype=np.uint8)
        pairs = np.zeros((N, 2), dtype=np.int32) if want_pairs else None
        counts = np.zeros(N, dtype=COUNTS_DTYPE) if want_counts else None
        self.L.ko_feasible_batch(self.h, Q.ctypes.data_as(C.POINTER(C.c_double)), N,
                                 out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 pairs.ctypes.data_as(C.POINTER(C.c_int32)) if want_pairs else None,
                                 counts.ctypes.data_as(C.c_void_p) if want_counts else None, int(nthreads))
        res = [out]
        if want_pairs:
            res.append(pairs)
        if want_counts:
            res.append(counts)
        return res[0] if len(res) == 1 else tuple(res)

    def cspace_distance(self, a, b, weights=None) -> float:
        a_, ap = _d(a)
        b_, bp = _d(b)
        w_, wp = (None, None) if weights is None else _d(weights)
        return float(self.L.ko_cspace_distance(self.h, ap, bp, wp))

    def interpolate(self, a, b, u) -> np.ndarray:
        a_, ap = _d(a)
        b_, bp = _d(b)
        out = np.zeros(self.nlinks)
        self.L.ko_interpolate(self.h, ap, bp, float(u), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def edge_visible(self, a, b, eps=0.01, weights=None):
        a_, ap = _d(a)
        b_, bp = _d(b)
        w_, wp = (None, None) if weights is None else _d(weights)
        n = C.c_int32(0)
        v = bool(self.L.ko_edge_visible(self.h, ap, bp, float(eps), wp, C.byref(n), None))
        return v, int(n.value)

    def edges_visible_batch(self, A, B, eps=0.01, weights=None, nthreads=0):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, self.nlinks)
        B = np.ascontiguousarray(B, dtype=np.float64).reshape(-1, self.nlinks)
        N = A.shape[0]
        out = np.zeros(N, dtype=np.uint8)
        nchecks = np.zeros(N, dtype=np.int32)
        w_, wp = (None, None) if weights is None else _d(weights)
        self.L.ko_edges_visible_batch(self.h, A.ctypes.data_as(C.POINTER(C.c_double)), B.ctypes.data_as(C.POINTER(C.c_double)),
                                      N, float(eps), wp, out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      nchecks.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
        return out, nchecks

    def distance(self, q, upper_bound=np.inf, include_self=False):
        q_, qp = _d(q)
        pr = np.zeros(2, dtype=np.int32)
        d = float(self.L.ko_distance(self.h, qp, float(upper_bound), int(include_self), pr.ctypes.data_as(C.POINTER(C.c_int32)), None))
        return d, (int(pr[0]), int(pr[1]))

    def distance_batch(self, Q, upper_bound=np.inf, include_self=False, nthreads=0):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nlinks)
        N = Q.shape[0]
        d = np.zeros(N)
        pr = np.zeros((N, 2), dtype=np.int32)
        self.L.ko_distance_batch(self.h, Q.ctypes.data_as(C.POINTER(C.c_double)), N, float(upper_bound), int(include_self),
                                 d.ctypes.data_as(C.POINTER(C.c_double)), pr.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
        return d, pr

    def geom_collides(self, ga, Ta, gb, Tb) -> bool:
        a_, ap = _d(Ta)
        b_, bp = _d(Tb)
        return bool(self.L.ko_geom_collides(self.h, ga, ap, gb, bp))

    def geom_within_distance(self, ga, Ta, gb, Tb, tol) -> bool:
        a_, ap = _d(Ta)
        b_, bp = _d(Tb)
        return bool(self.L.ko_geom_within_distance(self.h, ga, ap, gb, bp, float(tol)))

    def geom_distance(self, ga, Ta, gb, Tb, upper_bound=np.inf) -> float:
        a_, ap = _d(Ta)
        b_, bp = _d(Tb)
        return float(self.L.ko_geom_distance(self.h, ga, ap, gb, bp, float(upper_bound)))

    def geom_distance_brute(self, ga, Ta, gb, Tb) -> float:
        a_, ap = _d(Ta)
        b_, bp = _d(Tb)
        return float(self.L.ko_geom_distance_brute(self.h, ga, ap, gb, bp))

    def geom_penetration(self, ga, Ta, gb, Tb, tol=0.0) -> float:
        a_, ap = _d(Ta)
        b_, bp = _d(Tb)
        return float(self.L.ko_geom_penetration(self.h, ga, ap, gb, bp, float(tol)))

    def raycast_batch(self, q, rays, ignore_ids=None, nthreads=0):
        """WorldModel::RayCast / RayCastIgnore (World.cpp:465-588) of N rays (source xyz, direction xyz) with the robot at q (None: robot
        left out): (world id or -1, distance along the normalised direction or inf, element index)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        N = rays.shape[0]
        ids, dist, elem = np.zeros(N, dtype=np.int32), np.zeros(N), np.zeros(N, dtype=np.int32)
        q_, qp = (None, None) if q is None else _d(q)
        ig_, igp = (None, None) if ignore_ids is None else _u8(ignore_ids)
        self.L.ko_raycast_batch(self.h, qp, rays.ctypes.data_as(C.POINTER(C.c_double)), N, igp, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                dist.ctypes.data_as(C.POINTER(C.c_double)), elem.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
        return ids, dist, elem

    def raycast_counts(self, q, rays):
        """box / triangle / sphere tests of the per-body ray loop, averaged per ray (roofline of bench.py's ray workload)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        q_, qp = (None, None) if q is None else _d(q)
        c = Counts()
        self.L.ko_raycast_counts(self.h, qp, rays.ctypes.data_as(C.POINTER(C.c_double)), len(rays), C.byref(c))
        return {k: getattr(c, k) / max(1, len(rays)) for k in ("n_box", "n_node", "n_tri", "n_pt")}

    def geom_raycast(self, g, T, s, d, brute=False):
        """Geometry3D::rayCast_ext (Python/klampt/src/geometry.cpp:1837-1852): (hit, distance, element)"""
        T_, Tp = _d(T)
        s_, sp = _d(s)
        d_, dp_ = _d(d)
        dist, el = C.c_double(0), C.c_int32(-1)
        hit = self.L.ko_geom_raycast(self.h, int(g), Tp, sp, dp_, C.byref(dist), C.byref(el), int(brute))
        return bool(hit), float(dist.value), int(el.value)

    def geom_aabb(self, g, T):
        T_, Tp = _d(T)
        lo, hi = np.zeros(3), np.zeros(3)
        self.L.ko_geom_aabb(self.h, g, Tp, lo.ctypes.data_as(C.POINTER(C.c_double)), hi.ctypes.data_as(C.POINTER(C.c_double)))
        return lo, hi


def colliding_pairs(orc: "OracleWorld", q):
    """every enabled (idA, idB) world-id pair that collides at q, by explicit per-pair queries (test helper for
    kb_colliding_pairs_batch); None if the joint / driver limits fail"""
    spec = orc.spec
    if not orc.check_joint_limits(q):
        return None
    T = orc.fk(q)
    mask = orc.pair_mask()
    r = spec.robot
    out = set()
    I12 = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    env = [(spec.terrain_id(i), g, I12) for i, g in enumerate(spec.terrains)] + [(spec.rigid_object_id(i), g, Tm) for i, (g, Tm) in enumerate(spec.objects)]
    for j, gj in enumerate(r.link_geom):
        if gj < 0 or spec.geoms[gj].num_elements() == 0:
            continue
        lid = spec.robot_link_id(j)
        for (eid, ge, Te) in env:
            if ge < 0 or spec.geoms[ge].num_elements() == 0:
                continue
            if (mask[lid, eid] or mask[eid, lid]) and orc.geom_collides(gj, T[j], ge, Te):
                out.add((lid, eid))
        for k in range(j + 1, r.L):
            gk = r.link_geom[k]
            if gk < 0 or spec.geoms[gk].num_elements() == 0:
                continue
            kid = spec.robot_link_id(k)
            if (mask[lid, kid] or mask[lid, lid]) and orc.geom_collides(gj, T[j], gk, T[k]):
                out.add((lid, kid))
    return out
