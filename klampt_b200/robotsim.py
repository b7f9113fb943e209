"""Host-side mirror of the slice of Klamp't's ``robotsim`` API that sits on the feasibility hot path
(reference Python/klampt/src/robotmodel.h:141-191,603,850-855 and src/geometry.h:810-1108), with the same
names, argument meaning and error behaviour, so code written against ``klampt.WorldModel`` / ``RobotModel`` /
``Geometry3D`` for collision checking keeps working.  Every query below runs on the GPU through the C ABI
(klampt_b200.engine.Engine); the single-configuration calls are there for drop-in compatibility, the batch
entry points (``RobotModel.selfCollidesBatch``, ``klampt_b200.robotcspace.RobotCSpace.feasible_batch`` ...)
are what the hot path is for.

Conventions kept from the reference: rotations are column-major 9-lists (Python/klampt/math/so3.py:1-14),
``getBB`` returns (bmin, bmax), rigid-object transforms are pushed at query time, geometry margins add to
the collision threshold.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import so3
from .worldspec import (GeomSpec, RobotSpec, WorldSpec, REVOLUTE, PRISMATIC, JOINT_NORMAL, IDENTITY12)


class TriangleMesh:
    """vertices (n,3) float64, indices (m,3) int32 (Python/klampt/src/geometry.h:17-19)"""

    def __init__(self, vertices=None, indices=None):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.indices = np.zeros((0, 3), dtype=np.int32) if indices is None else np.asarray(indices, dtype=np.int32).reshape(-1, 3)


class PointCloud:
    """points (n,3) float64 (+ optional per-point 'radius' property; src/geometry.h:140-148)"""

    def __init__(self, points=None, radius=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64).reshape(-1, 3)
        self.radius = None if radius is None else np.asarray(radius, dtype=np.float64).reshape(-1)


class GeometricPrimitive:
    def __init__(self, type: str = "", properties: Sequence[float] = ()):
        self.type = type
        self.properties = list(properties)

    def setPoint(self, p):
        self.type, self.properties = "Point", [float(x) for x in p]

    def setSphere(self, c, r):
        self.type, self.properties = "Sphere", [float(x) for x in c] + [float(r)]

    def setAABB(self, bmin, bmax):
        self.type, self.properties = "AABB", [float(x) for x in bmin] + [float(x) for x in bmax]

    def setBox(self, ori, R, dims):
        """Klamp't's convention (Python/klampt/src/geometry.h GeometricPrimitive.setBox): `ori` is the box's ORIGIN corner, `R`
        a column-major so3 9-list whose columns are the box axes, `dims` the full edge lengths"""
        self.type, self.properties = "Box", [float(x) for x in ori] + [float(x) for x in R] + [float(x) for x in dims]

    def setSegment(self, a, b):
        self.type, self.properties = "Segment", [float(x) for x in a] + [float(x) for x in b]

    def setTriangle(self, a, b, c):
        self.type, self.properties = "Triangle", [float(x) for x in a] + [float(x) for x in b] + [float(x) for x in c]


class DistanceQuerySettings:
    def __init__(self):
        self.relErr = 0.0
        self.absErr = 0.0
        self.upperBound = float("inf")


class DistanceQueryResult:
    """src/geometry.h:631-694: d, closest points (world frame, on the margin-inflated surfaces) and element indices; gradients:
    the unit direction between the closest points when they are distinct (grad1 = d d / d cp1 = -(cp2 - cp1) / |..|, grad2 = -grad1)"""

    def __init__(self, d, cp1=None, cp2=None, elem1=-1, elem2=-1):
        self.d = d
        self.hasClosestPoints = cp1 is not None and not any(x != x for x in cp1)
        self.cp1 = [float(x) for x in cp1] if self.hasClosestPoints else []
        self.cp2 = [float(x) for x in cp2] if self.hasClosestPoints else []
        self.elem1, self.elem2 = (int(elem1), int(elem2)) if self.hasClosestPoints else (-1, -1)
        self.hasGradients = False
        self.grad1 = self.grad2 = []
        if self.hasClosestPoints:
            v = np.asarray(self.cp2) - np.asarray(self.cp1)
            n = float(np.linalg.norm(v))
            if n > 0 and d > 0:
                self.hasGradients = True
                self.grad1, self.grad2 = list(-v / n), list(v / n)


_PAIR_ENGINES = {}


class Geometry3D:
    """AnyCollisionGeometry3D: data + current transform + collision margin."""

    def __init__(self, data=None):
        self._kind = "empty"
        self._data = None
        self._R = so3.identity()
        self._t = [0.0, 0.0, 0.0]
        self._margin = 0.0
        self._version = 0
        if isinstance(data, TriangleMesh):
            self.setTriangleMesh(data)
        elif isinstance(data, PointCloud):
            self.setPointCloud(data)
        elif isinstance(data, GeometricPrimitive):
            self.setGeometricPrimitive(data)

    # ---- data
    def type(self) -> str:
        return {"empty": "", "mesh": "TriangleMesh", "cloud": "PointCloud", "prim": "GeometricPrimitive"}[self._kind]

    def empty(self) -> bool:
        return self._kind == "empty" or self.numElements() == 0

    def numElements(self) -> int:
        if self._kind == "mesh":
            return len(self._data.indices)
        if self._kind == "cloud":
            return len(self._data.points)
        return 1 if self._kind == "prim" else 0

    def setTriangleMesh(self, m: TriangleMesh):
        self._kind, self._data, self._version = "mesh", m, self._version + 1

    def setPointCloud(self, pc: PointCloud):
        self._kind, self._data, self._version = "cloud", pc, self._version + 1

    def setGeometricPrimitive(self, p: GeometricPrimitive):
        if p.type not in ("Point", "Sphere", "Segment", "Triangle", "AABB", "Box"):
            raise ValueError("GeometricPrimitive type %r is not supported by the batched engine (Point, Sphere, Segment, Triangle, AABB and Box are)" % p.type)
        self._kind, self._data, self._version = "prim", p, self._version + 1

    def getTriangleMesh(self) -> TriangleMesh:
        if self._kind != "mesh":
            raise RuntimeError("Geometry is not a TriangleMesh")
        return self._data

    def getPointCloud(self) -> PointCloud:
        if self._kind != "cloud":
            raise RuntimeError("Geometry is not a PointCloud")
        return self._data

    def getGeometricPrimitive(self) -> GeometricPrimitive:
        if self._kind != "prim":
            raise RuntimeError("Geometry is not a GeometricPrimitive")
        return self._data

    # ---- transform / margin
    def setCurrentTransform(self, R: Sequence[float], t: Sequence[float]):
        self._R, self._t = list(R), list(t)

    def getCurrentTransform(self) -> Tuple[List[float], List[float]]:
        return list(self._R), list(self._t)

    def setCollisionMargin(self, margin: float):
        if margin < 0:
            raise ValueError("margin must be >= 0")
        self._margin, self._version = float(margin), self._version + 1

    def getCollisionMargin(self) -> float:
        return self._margin

    def _T12(self) -> np.ndarray:
        return so3.to_rowmajor12(self._R, self._t)

    def to_spec(self) -> GeomSpec:
        if self._kind == "mesh":
            return GeomSpec.mesh(self._data.vertices, self._data.indices, self._margin)
        if self._kind == "cloud":
            return GeomSpec.cloud(self._data.points, self._data.radius, self._margin)
        if self._kind == "prim":
            p = self._data
            if p.type == "Segment":
                return GeomSpec.segment(p.properties[:3], p.properties[3:6], self._margin)
            if p.type == "Triangle":
                return GeomSpec.triangle(p.properties[:3], p.properties[3:6], p.properties[6:9], self._margin)
            if p.type == "AABB":
                return GeomSpec.aabb(p.properties[:3], p.properties[3:6], self._margin)
            if p.type == "Box":
                ori, M, dims = np.asarray(p.properties[:3]), so3.matrix(p.properties[3:12]), np.asarray(p.properties[12:15])
                return GeomSpec.box(ori + M @ (0.5 * dims), M, 0.5 * dims, self._margin)
            return GeomSpec.point(p.properties[:3], self._margin) if p.type == "Point" else GeomSpec.sphere(p.properties[:3], p.properties[3], self._margin)
        return GeomSpec("empty")

    def _local_points(self) -> np.ndarray:
        if self._kind == "mesh":
            return self._data.vertices
        if self._kind == "cloud":
            return self._data.points
        if self._kind == "prim":
            if self._data.type in ("AABB", "Box"):
                g = self.to_spec()
                c, M, h = g.params[:3], g.params[3:12].reshape(3, 3), g.params[12:15]
                corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64) * h
                return corners @ M.T + c
            n = {"Triangle": 9, "Segment": 6}.get(self._data.type, 3)
            return np.asarray(self._data.properties[:n], dtype=np.float64).reshape(-1, 3)
        return np.zeros((0, 3))

    def getBB(self):
        """O(1)-style loose box: the local AABB carried through the current transform and grown by the margin
        (AnyCollisionGeometry3D::GetAABB as used by Cpp/Planning/PlannerSettings.cpp:250-254)."""
        return self._bb(tight=False)

    def getBBTight(self):
        return self._bb(tight=True)

    def _bb(self, tight: bool):
        p = self._local_points()
        if len(p) == 0:
            return [float("inf")] * 3, [float("-inf")] * 3
        r = 0.0
        if self._kind == "cloud" and self._data.radius is not None:
            r = float(np.max(self._data.radius))
        if self._kind == "prim" and self._data.type == "Sphere":
            r = float(self._data.properties[3])
        M, t = so3.matrix(self._R), np.asarray(self._t)
        if tight:
            w = p @ M.T + t
            lo, hi = w.min(axis=0), w.max(axis=0)
        else:
            lo0, hi0 = p.min(axis=0), p.max(axis=0)
            c, h = 0.5 * (lo0 + hi0), 0.5 * (hi0 - lo0)
            wc, e = M @ c + t, np.abs(M) @ h
            lo, hi = wc - e, wc + e
        g = r + self._margin
        return list(lo - g), list(hi + g)

    # ---- pair queries (GPU)
    def _pair_engine(self, other: "Geometry3D"):
        from .engine import Engine
        key = (id(self), self._version, id(other), other._version)
        hit = _PAIR_ENGINES.get(key)
        if hit is None:
            if len(_PAIR_ENGINES) > 64:
                _PAIR_ENGINES.clear()
            w = WorldSpec()
            ga, gb = w.add_geom(self.to_spec()), w.add_geom(other.to_spec())
            w.robot = RobotSpec(parents=np.array([-1], dtype=np.int32), linktype=np.array([REVOLUTE], dtype=np.uint8),
                                axis=np.array([[0.0, 0.0, 1.0]]), T0=IDENTITY12.reshape(1, 12).copy(), qmin=np.zeros(1), qmax=np.zeros(1),
                                link_geom=[-1])
            hit = (Engine(w), ga, gb)
            _PAIR_ENGINES[key] = hit
        return hit

    def collides(self, other: "Geometry3D") -> bool:
        if self.empty() or other.empty():
            return False
        eng, ga, gb = self._pair_engine(other)
        return bool(eng.geom_collides_batch(ga, self._T12(), gb, other._T12())[0])

    def withinDistance(self, other: "Geometry3D", tol: float) -> bool:
        if self.empty() or other.empty():
            return False
        eng, ga, gb = self._pair_engine(other)
        return bool(eng.geom_collides_batch(ga, self._T12(), gb, other._T12(), tol=tol)[0])

    def distance_simple(self, other: "Geometry3D", relErr: float = 0, absErr: float = 0) -> float:
        st = DistanceQuerySettings()
        st.relErr, st.absErr = relErr, absErr
        return self.distance_ext(other, st).d

    def distance(self, other: "Geometry3D") -> DistanceQueryResult:
        return self.distance_ext(other, DistanceQuerySettings())

    def distance_point(self, pt) -> DistanceQueryResult:
        """Geometry3D.distance_point (reference Python/klampt/src/geometry.h:1011-1030): distance from this geometry, at its
        current transform, to a world-space point (margins subtracted)."""
        return self.distance_point_ext(pt, DistanceQuerySettings())

    def distance_point_ext(self, pt, settings: DistanceQuerySettings) -> DistanceQueryResult:
        if self.empty():
            raise RuntimeError("Distance queries not implemented yet for those types of geometry, or geometries are content-empty?")
        eng, ga, gb = self._pair_engine(_POINT_PROBE)
        Tb = IDENTITY12.copy()
        Tb[9:12] = np.asarray(pt, dtype=np.float64)
        d, cp, el = eng.geom_distance_batch_ex(ga, self._T12(), gb, Tb, upper_bound=settings.upperBound, abs_err=settings.absErr, rel_err=settings.relErr)
        return DistanceQueryResult(float(d[0]), cp[0, 0], cp[0, 1], el[0, 0], el[0, 1])

    def distance_points_batch(self, pts, upper_bound: float = float("inf")) -> np.ndarray:
        """distance_point for N world-space points in one launch: a point primitive at N translations"""
        if self.empty():
            raise RuntimeError("Distance queries not implemented yet for those types of geometry, or geometries are content-empty?")
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        probe = _POINT_PROBE
        eng, ga, gb = self._pair_engine(probe)
        Ta = np.tile(self._T12(), (len(pts), 1))
        Tb = np.tile(IDENTITY12, (len(pts), 1))
        Tb[:, 9:12] = pts
        return eng.geom_distance_batch(ga, Ta, gb, Tb, upper_bound=upper_bound)

    def distance_ext(self, other: "Geometry3D", settings: DistanceQuerySettings) -> DistanceQueryResult:
        if self.empty() or other.empty():
            raise RuntimeError("Distance queries not implemented yet for those types of geometry, or geometries are content-empty?")
        eng, ga, gb = self._pair_engine(other)
        d, cp, el = eng.geom_distance_batch_ex(ga, self._T12(), gb, other._T12(), upper_bound=settings.upperBound,
                                               abs_err=settings.absErr, rel_err=settings.relErr)
        return DistanceQueryResult(float(d[0]), cp[0, 0], cp[0, 1], el[0, 0], el[0, 1])


    # ---- ray casts (GPU; reference Python/klampt/src/geometry.h:1106-1140, src/geometry.cpp:1821-1852)
    def rayCastBatch(self, rays):
        """rayCast_ext for N rays (rows of source xyz, direction xyz) in one launch: (element index or -1, distance along the
        normalised direction or inf).  The hit point is source + distance * direction / |direction|."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        if self.empty():
            return np.full(len(rays), -1, dtype=np.int32), np.full(len(rays), np.inf)
        eng, ga, _ = self._pair_engine(_POINT_PROBE)
        return eng.geom_raycast_batch(ga, self._T12(), rays)

    def rayCast_ext(self, s, d):
        """(hit_element, pt): the element hit (-1: none) and the hit point in world coordinates"""
        s, d = np.asarray(s, dtype=np.float64), np.asarray(d, dtype=np.float64)
        el, dist = self.rayCastBatch(np.concatenate([s, d])[None, :])
        if el[0] < 0:
            return -1, [0.0, 0.0, 0.0]
        return int(el[0]), list(s + dist[0] * d / np.linalg.norm(d))

    def rayCast(self, s, d):
        """(hit, pt): whether the ray from s along d hits the geometry, and where (world coordinates)"""
        el, pt = self.rayCast_ext(s, d)
        return el >= 0, pt


def _prim_from_spec(g: GeomSpec) -> GeometricPrimitive:
    if g.kind == "box":
        c, M, h = g.params[:3], g.params[3:12].reshape(3, 3), g.params[12:15]
        p = GeometricPrimitive()
        p.setBox(list(c - M @ h), so3.from_matrix(M), list(2.0 * h))
        return p
    return GeometricPrimitive({"sphere": "Sphere", "point": "Point", "triangle": "Triangle", "segment": "Segment"}[g.kind], list(g.params))


_POINT_PROBE = Geometry3D()
_POINT_PROBE.setGeometricPrimitive(GeometricPrimitive("Point", [0.0, 0.0, 0.0]))


class _Named:
    def __init__(self, name):
        self._name = name

    def getName(self) -> str:
        return self._name

    def setName(self, n: str):
        self._name = n


class TerrainModel(_Named):
    def __init__(self, world, index, name):
        super().__init__(name)
        self.world, self.index = world, index
        self._geom = Geometry3D()

    def getID(self) -> int:
        """world ID (Cpp/Modeling/World.cpp:47-53: terrains, then rigid objects, then per robot its ID and its links)"""
        return self.world.terrainID(self.index)

    def geometry(self) -> Geometry3D:
        return self._geom


class RigidObjectModel(_Named):
    def __init__(self, world, index, name):
        super().__init__(name)
        self.world, self.index = world, index
        self._geom = Geometry3D()

    def getID(self) -> int:
        return self.world.rigidObjectID(self.index)

    def geometry(self) -> Geometry3D:
        return self._geom

    def setTransform(self, R, t):
        self._geom.setCurrentTransform(R, t)

    def getTransform(self):
        return self._geom.getCurrentTransform()


class RobotModelLink(_Named):
    def __init__(self, robot: "RobotModel", index: int, name: str):
        super().__init__(name)
        self._robot, self.index = robot, index
        self.world = getattr(robot, "world", None)
        self._geom = Geometry3D()

    def robot(self) -> "RobotModel":
        return self._robot

    def getID(self) -> int:
        return self.world.robotLinkID(self._robot.index, self.index)

    def getIndex(self) -> int:
        return self.index

    def getParent(self) -> int:
        return int(self._robot._parents[self.index])

    def getAxis(self) -> List[float]:
        return list(self._robot._axis[self.index])

    def isPrismatic(self) -> bool:
        return int(self._robot._linktype[self.index]) == PRISMATIC

    def isRevolute(self) -> bool:
        return not self.isPrismatic()

    def getParentTransform(self):
        return so3.from_rowmajor12(self._robot._T0[self.index])

    def setParentTransform(self, R, t):
        self._robot._T0[self.index] = so3.to_rowmajor12(R, t)
        self._robot._dirty()

    def geometry(self) -> Geometry3D:
        return self._geom

    def getTransform(self):
        return self._geom.getCurrentTransform()


class RobotModelDriver:
    """the part of RobotModelDriver the planning layer reads (Python/klampt/src/robotmodel.h; plan/cspaceutils.py:335-352): a normal
    driver moves one link; an affine driver moves several with q_link = scale * value + offset"""

    def __init__(self, robot: "RobotModel", index: int, spec):
        self._robot, self.index, self._spec = robot, index, spec

    def robot(self) -> "RobotModel":
        return self._robot

    def getType(self) -> str:
        return "affine" if len(self._spec.links) > 1 else "normal"

    def getAffectedLink(self) -> int:
        return int(self._spec.links[0])

    def getAffectedLinks(self) -> List[int]:
        return [int(k) for k in self._spec.links]

    def getAffineCoeffs(self):
        return [float(s) for s in self._spec.scale], [float(o) for o in self._spec.offset]

    def getLimits(self):
        return [float(self._spec.qmin), float(self._spec.qmax)]


class RobotModel(_Named):
    """RobotKinematics3D + RobotWithGeometry data model with the collision-relevant robotsim methods."""

    def __init__(self, world, index, name, spec: Optional[RobotSpec] = None, geoms: Optional[List[GeomSpec]] = None):
        super().__init__(name)
        self.world, self.index = world, index
        self._links: List[RobotModelLink] = []
        self._engine = None
        self._nl_engine = None
        self._q = np.zeros(0)
        if spec is not None:
            self._from_spec(spec, geoms)

    def getID(self) -> int:
        return self.world.robotID(self.index)

    def _from_spec(self, spec: RobotSpec, geoms):
        L = spec.L
        self._parents = np.array(spec.parents, dtype=np.int32)
        self._linktype = np.array(spec.linktype, dtype=np.uint8)
        self._axis = np.array(spec.axis, dtype=np.float64).reshape(L, 3)
        self._T0 = np.array(spec.T0, dtype=np.float64).reshape(L, 12)
        self._qmin, self._qmax = np.array(spec.qmin, dtype=np.float64), np.array(spec.qmax, dtype=np.float64)
        self._joint_type = np.full(L, JOINT_NORMAL, dtype=np.uint8) if spec.joint_type is None else np.array(spec.joint_type, dtype=np.uint8)
        self._joint_link = np.arange(L, dtype=np.int32) if spec.joint_link is None else np.array(spec.joint_link, dtype=np.int32)
        self._joint_base = None if getattr(spec, "joint_base", None) is None else np.array(spec.joint_base, dtype=np.int32)
        self._drivers = list(spec.drivers)
        self._self_edits = list(spec.self_collision_edits)
        names = spec.names or ["Link_%d" % i for i in range(L)]
        self._links = [RobotModelLink(self, i, names[i]) for i in range(L)]
        for i, gi in enumerate(spec.link_geom):
            if gi >= 0 and geoms is not None:
                g = geoms[gi]
                if g.kind == "mesh":
                    self._links[i]._geom.setTriangleMesh(TriangleMesh(g.verts, g.tris))
                elif g.kind == "cloud":
                    self._links[i]._geom.setPointCloud(PointCloud(g.points, g.radius))
                elif g.kind in ("sphere", "point", "segment", "triangle", "box"):
                    self._links[i]._geom.setGeometricPrimitive(_prim_from_spec(g))
                self._links[i]._geom._margin = g.margin
        self._q = np.clip(np.zeros(L), self._qmin, self._qmax)
        self._selfcol = None

    def _dirty(self):
        self._engine = None
        self._nl_engine = None
        if self.world is not None:
            self.world._dirty()

    # ---- structure
    def numDrivers(self) -> int:
        return len(self._drivers)

    def driver(self, i: int) -> RobotModelDriver:
        return RobotModelDriver(self, i, self._drivers[i])

    def numLinks(self) -> int:
        return len(self._links)

    def link(self, i) -> RobotModelLink:
        if isinstance(i, str):
            for l in self._links:
                if l.getName() == i:
                    return l
            raise KeyError(i)
        return self._links[i]

    def getJointLimits(self):
        return list(self._qmin), list(self._qmax)

    def setJointLimits(self, qmin, qmax):
        self._qmin, self._qmax = np.asarray(qmin, dtype=np.float64), np.asarray(qmax, dtype=np.float64)
        self._dirty()

    # ---- self-collision pair set (robotsim.cpp:5504-5527)
    def _selfcol_matrix(self) -> np.ndarray:
        if self._selfcol is None:
            L = self.numLinks()
            m = np.zeros((L, L), dtype=bool)
            for i in range(L):
                for j in range(i + 1, L):
                    m[i, j] = (not self._links[i]._geom.empty() and not self._links[j]._geom.empty()
                               and self._parents[j] != i and self._parents[i] != j)
            for (i, j, en) in self._self_edits:
                a, b = min(i, j), max(i, j)
                m[a, b] = bool(en) and not self._links[a]._geom.empty() and not self._links[b]._geom.empty()
            self._selfcol = m
        return self._selfcol

    def selfCollisionEnabled(self, link1: int, link2: int) -> bool:
        a, b = min(link1, link2), max(link1, link2)
        return False if a == b else bool(self._selfcol_matrix()[a, b])

    def enableSelfCollision(self, link1: int, link2: int, value: bool):
        if link1 == link2:
            return
        self._self_edits.append((min(link1, link2), max(link1, link2), bool(value)))
        self._selfcol = None
        self._dirty()

    # ---- spec / engine
    def to_spec(self, world: WorldSpec) -> RobotSpec:
        link_geom = []
        for l in self._links:
            link_geom.append(-1 if l._geom.empty() else world.add_geom(l._geom.to_spec()))
        return RobotSpec(parents=self._parents.copy(), linktype=self._linktype.copy(), axis=self._axis.copy(), T0=self._T0.copy(),
                         qmin=self._qmin.copy(), qmax=self._qmax.copy(), link_geom=link_geom, joint_type=self._joint_type.copy(),
                         joint_link=self._joint_link.copy(), joint_base=None if self._joint_base is None else self._joint_base.copy(),
                         drivers=list(self._drivers), self_collision_edits=list(self._self_edits),
                         names=[l.getName() for l in self._links])

    def _self_engine(self):
        """engine holding only this robot: FK and self-collision queries"""
        if self._engine is None:
            from .engine import Engine
            w = WorldSpec()
            w.robot = self.to_spec(w)
            self._engine = Engine(w)
        return self._engine

    # ---- configuration (robotsim.cpp:5311-5321: q copy, UpdateFrames, UpdateGeometry)
    def getConfig(self) -> List[float]:
        return list(self._q)

    def setConfig(self, q: Sequence[float]):
        if len(q) != self.numLinks():
            raise ValueError("Invalid size of configuration, %d != %d" % (len(q), self.numLinks()))
        self._q = np.asarray(q, dtype=np.float64).copy()
        T = self._self_engine().fk_batch(self._q)[0]
        for i, l in enumerate(self._links):
            R, t = so3.from_rowmajor12(T[i])
            l._geom.setCurrentTransform(R, t)

    def selfCollides(self) -> bool:
        """RobotWithGeometry::SelfCollision at the current configuration (robotsim.cpp:5529-5539)"""
        return bool(self.selfCollidesBatch(self._q.reshape(1, -1))[0])

    def selfCollidesBatch(self, Q) -> np.ndarray:
        """batch form: True where any enabled self pair collides.  Joint limits are not part of this query, so
        limit-violating rows are checked with limits widened to +-inf."""
        eng = self._nolimit_engine()
        return eng.feasible_batch(Q) == 0

    def _nolimit_engine(self):
        if self._nl_engine is None:
            from .engine import Engine
            w = WorldSpec()
            spec = self.to_spec(w)
            spec.qmin = np.full(spec.L, -np.inf)
            spec.qmax = np.full(spec.L, np.inf)
            spec.drivers = []
            w.robot = spec
            self._nl_engine = Engine(w)
        return self._nl_engine

    # ---- C-space helpers (RobotModel.interpolate / distance -> Klampt::Interpolate / Distance)
    def _joint_indices(self, j: int) -> List[int]:
        """RobotModel::GetJointIndices (reference Cpp/Modeling/Robot.cpp:2120-2144): links a joint drives, root to tip"""
        link = int(self._joint_link[j])
        if self._joint_type[j] in (0, 1, 2) or self._joint_base is None:
            return [link]
        out = []
        while link != int(self._joint_base[j]):
            out.append(link)
            link = int(self._parents[link])
        return out[::-1]

    @staticmethod
    def _short_arc(x, y):
        x, y = x % (2 * np.pi), y % (2 * np.pi)
        d = y - x
        return x, d - 2 * np.pi if d > np.pi else (d + 2 * np.pi if d < -np.pi else d)

    def interpolate(self, a, b, u) -> List[float]:
        """Klampt::Interpolate (reference Cpp/Modeling/Interpolate.cpp:10-71)"""
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        out = a * (1.0 - u)
        out += b * u
        for j, (jt, k) in enumerate(zip(self._joint_type, self._joint_link)):
            if jt == 2 or jt == 4:   # Spin / FloatingPlanar angle: shortest arc
                if jt == 4:
                    k = self._joint_indices(j)[2]
                x, d = self._short_arc(a[k], b[k])
                out[k] = (x + u * d) % (2 * np.pi)
            elif jt == 3 or jt == 5:   # Floating / BallAndSocket: Euler ZYX triplet along the SO(3) geodesic
                ix = self._joint_indices(j)[-3:]
                out[ix] = so3.euler_zyx_interp(a[ix], b[ix], u)
        return list(out)

    def distance(self, a, b) -> float:
        """Klampt::Distance, norm 2, floatingRotationWeight 1 (reference Cpp/Modeling/Interpolate.cpp:208-343)"""
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        s = 0.0
        for j, (jt, k) in enumerate(zip(self._joint_type, self._joint_link)):
            if jt == 1:
                s += (a[k] - b[k]) ** 2
            elif jt == 2:
                _, d = self._short_arc(b[k], a[k])
                s += d * d
            elif jt == 3 or jt == 5:
                ix = self._joint_indices(j)
                if jt == 3:
                    s += float(((a[ix[:3]] - b[ix[:3]]) ** 2).sum())
                s += so3.euler_zyx_angle_between(a[ix[-3:]], b[ix[-3:]]) ** 2
        return float(np.sqrt(s))


class WorldModel:
    """terrains / rigid objects / robots with the reference's ID scheme (Cpp/Modeling/World.cpp:47-53,110-180)."""

    def __init__(self):
        self._terrains: List[TerrainModel] = []
        self._objects: List[RigidObjectModel] = []
        self._robots: List[RobotModel] = []
        self._version = 0

    def _dirty(self):
        self._version += 1

    @staticmethod
    def from_spec(spec: WorldSpec) -> "WorldModel":
        w = WorldModel()
        for i, gi in enumerate(spec.terrains):
            t = w.makeTerrain("terrain%d" % i)
            _fill(t._geom, spec.geoms[gi] if gi >= 0 else None)
        for i, (gi, T) in enumerate(spec.objects):
            o = w.makeRigidObject("object%d" % i)
            _fill(o._geom, spec.geoms[gi] if gi >= 0 else None)
            o.setTransform(*so3.from_rowmajor12(T))
        if spec.robot is not None:
            w.addRobot("robot", spec.robot, spec.geoms)
        return w

    def makeTerrain(self, name: str) -> TerrainModel:
        self._terrains.append(TerrainModel(self, len(self._terrains), name))
        self._dirty()
        return self._terrains[-1]

    def makeRigidObject(self, name: str) -> RigidObjectModel:
        self._objects.append(RigidObjectModel(self, len(self._objects), name))
        self._dirty()
        return self._objects[-1]

    def addRobot(self, name: str, spec: RobotSpec, geoms: List[GeomSpec]) -> RobotModel:
        self._robots.append(RobotModel(self, len(self._robots), name, spec, geoms))
        self._dirty()
        return self._robots[-1]

    def numTerrains(self):
        return len(self._terrains)

    def numRigidObjects(self):
        return len(self._objects)

    def numRobots(self):
        return len(self._robots)

    def terrain(self, i) -> TerrainModel:
        return self._terrains[i]

    def rigidObject(self, i) -> RigidObjectModel:
        return self._objects[i]

    def robot(self, i) -> RobotModel:
        return self._robots[i]

    def enableInitCollisions(self, enabled: bool):
        pass   # collision structures are always built at engine finalisation

    def numIDs(self) -> int:
        return len(self._terrains) + len(self._objects) + sum(1 + r.numLinks() for r in self._robots)

    def terrainID(self, i):
        return i

    def rigidObjectID(self, i):
        return len(self._terrains) + i

    def robotID(self, r=0):
        base = len(self._terrains) + len(self._objects)
        for k in range(r):
            base += 1 + self._robots[k].numLinks()
        return base

    def robotLinkID(self, r, j):
        return self.robotID(r) + 1 + j

    def to_spec(self, robot_index: int = 0, pair_mask: Optional[np.ndarray] = None) -> WorldSpec:
        """the world the engine works on: ONE active robot (SingleRobotCSpace holds one), the terrains, the rigid objects, and -- as
        SingleRobotCSpace::CheckCollisionFree checks the robot against "all other robots" too (RobotCSpace.cpp:794-823) -- the links of
        every other robot as rigid bodies at their current transforms.  The engine numbers those links after the rigid objects;
        ``spec.world_ids[engine id]`` gives the id of the reference's scheme (World.cpp:47-53: terrains, rigid objects, then per
        robot its id and its links), and ``pair_mask`` (reference numbering) is permuted accordingly."""
        if not self._robots:
            raise NotImplementedError("the batched engine needs a robot")
        w = WorldSpec()
        ids: List[int] = []
        for i, t in enumerate(self._terrains):
            w.terrains.append(-1 if t._geom.empty() else w.add_geom(t._geom.to_spec()))
            ids.append(self.terrainID(i))
        for i, o in enumerate(self._objects):
            gi = -1 if o._geom.empty() else w.add_geom(o._geom.to_spec())
            w.objects.append((gi, o._geom._T12()))
            ids.append(self.rigidObjectID(i))
        for r, rob in enumerate(self._robots):
            if r == robot_index:
                continue
            for j in range(rob.numLinks()):
                g = rob.link(j)._geom
                w.objects.append((-1 if g.empty() else w.add_geom(g.to_spec()), g._T12()))
                ids.append(self.robotLinkID(r, j))
        active = self._robots[robot_index]
        w.robot = active.to_spec(w)
        ids.append(self.robotID(robot_index))
        ids += [self.robotLinkID(robot_index, j) for j in range(active.numLinks())]
        multi = len(self._robots) > 1
        w.world_ids = np.asarray(ids, dtype=np.int32) if multi else None
        if pair_mask is not None and multi:
            pm = np.asarray(pair_mask, dtype=np.uint8)
            pair_mask = np.ascontiguousarray(pm[np.ix_(ids, ids)])
        w.pair_mask = pair_mask
        return w


def _fill(geom: Geometry3D, g: Optional[GeomSpec]):
    if g is None or g.kind == "empty":
        return
    if g.kind == "mesh":
        geom.setTriangleMesh(TriangleMesh(g.verts, g.tris))
    elif g.kind == "cloud":
        geom.setPointCloud(PointCloud(g.points, g.radius))
    else:
        geom.setGeometricPrimitive(_prim_from_spec(g))
    geom._margin = g.margin
