"""Plain-data description of one planning world (static geometry + one active robot).

This is the neutral container every front end produces (the robotsim-style model classes in
``klampt_b200.robotsim``, the ``.rob`` loader, the synthetic fixtures in ``klampt_b200.synth``) and
that the engine (``klampt_b200.engine.Engine``) uploads through the C ABI.

Reference semantics carried by the fields (paths relative to /root/reference):
  * world ID scheme: terrains ``[0,T)``, rigid objects ``[T,T+O)``, robot id, then L link ids
    (Cpp/Modeling/World.cpp:47-53,110-180);
  * transforms are 12 doubles, row-major R then t (file-format convention,
    Cpp/docs/Manual-FileTypes.md:35-37);
  * link recurrence ``T_World[i] = T_World[parent] * T0_Parent[i] * T_loc(q_i)``
    (Cpp/docs/Manual-Modeling.md:94,107-117);
  * collision margin adds to the collision threshold and subtracts from distances
    (Cpp/docs/Manual-Geometry.md:17).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

REVOLUTE, PRISMATIC = 0, 1
JOINT_WELD, JOINT_NORMAL, JOINT_SPIN, JOINT_FLOATING, JOINT_FLOATINGPLANAR, JOINT_BALLANDSOCKET, JOINT_CLOSED = range(7)
PRIM_POINT, PRIM_SPHERE, PRIM_TRIANGLE, PRIM_BOX, PRIM_AABB = 0, 1, 2, 3, 4

IDENTITY12 = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


@dataclass
class GeomSpec:
    """One collision geometry in its local frame (AnyCollisionGeometry3D minus the current transform)."""
    kind: str = "empty"                      # 'mesh' | 'cloud' | 'dyncloud' | 'sphere' | 'point' | 'segment' | 'triangle' | 'box' | 'empty'
    verts: Optional[np.ndarray] = None       # (nv,3) f64   (mesh)
    tris: Optional[np.ndarray] = None        # (nt,3) i32   (mesh)
    points: Optional[np.ndarray] = None      # (n,3)  f64   (cloud)
    radius: Optional[np.ndarray] = None      # (n,)   f64 or None (cloud)
    params: Optional[np.ndarray] = None      # sphere: cx,cy,cz,r ; point: x,y,z ; triangle: a,b,c (9) ; box: centre(3), row-major R whose columns are the axes (9), half dims(3)
    margin: float = 0.0

    @staticmethod
    def mesh(verts, tris, margin=0.0) -> "GeomSpec":
        return GeomSpec("mesh", verts=np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3),
                        tris=np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3), margin=float(margin))

    @staticmethod
    def cloud(points, radius=None, margin=0.0) -> "GeomSpec":
        r = None if radius is None else np.ascontiguousarray(radius, dtype=np.float64).reshape(-1)
        return GeomSpec("cloud", points=np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3), radius=r,
                        margin=float(margin))

    @staticmethod
    def dynamic_cloud(capacity: int, radius: float = 0.0, margin: float = 0.0) -> "GeomSpec":
        """a point cloud whose points are replaced between batches (Engine.update_pointcloud); starts empty"""
        return GeomSpec("dyncloud", params=np.array([float(capacity), float(radius)]), margin=float(margin))

    @staticmethod
    def sphere(center, r, margin=0.0) -> "GeomSpec":
        return GeomSpec("sphere", params=np.array([center[0], center[1], center[2], r], dtype=np.float64), margin=float(margin))

    @staticmethod
    def point(p, margin=0.0) -> "GeomSpec":
        return GeomSpec("point", params=np.array([p[0], p[1], p[2]], dtype=np.float64), margin=float(margin))

    @staticmethod
    def segment(a, b, margin=0.0) -> "GeomSpec":
        """Segment3D a-b (a != b)"""
        return GeomSpec("segment", params=np.concatenate([np.asarray(a, dtype=np.float64).reshape(3), np.asarray(b, dtype=np.float64).reshape(3)]),
                        margin=float(margin))

    @staticmethod
    def triangle(a, b, c, margin=0.0) -> "GeomSpec":
        return GeomSpec("triangle", params=np.concatenate([np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64),
                                                            np.asarray(c, dtype=np.float64)]), margin=float(margin))

    @staticmethod
    def box(center, R, half, margin=0.0) -> "GeomSpec":
        """solid oriented box (GeometricPrimitive3D Box3D): centre, 3x3 whose COLUMNS are the box axes, half dimensions"""
        return GeomSpec("box", params=np.concatenate([np.asarray(center, dtype=np.float64).reshape(3), np.asarray(R, dtype=np.float64).reshape(9),
                                                       np.asarray(half, dtype=np.float64).reshape(3)]), margin=float(margin))

    @staticmethod
    def aabb(lo, hi, margin=0.0) -> "GeomSpec":
        """solid axis-aligned box (AABB3D) in the geometry's local frame"""
        lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        return GeomSpec.box(0.5 * (lo + hi), np.eye(3), 0.5 * (hi - lo), margin)

    def num_elements(self) -> int:
        if self.kind == "mesh":
            return int(self.tris.shape[0])
        if self.kind == "cloud":
            return int(self.points.shape[0])
        if self.kind == "dyncloud":
            return int(self.params[0])
        return 0 if self.kind == "empty" else 1


@dataclass
class DriverSpec:
    """RobotModelDriver restricted to what CheckJointLimits reads (Cpp/Modeling/Robot.cpp:2166-2187):
    value = mean_k (q[link_k] - offset_k) / scale_k ; a Normal driver is the 1-link, scale 1, offset 0 case."""
    links: List[int]
    scale: List[float]
    offset: List[float]
    qmin: float
    qmax: float


@dataclass
class RobotSpec:
    parents: np.ndarray                      # (L,) i32, parents[i] < i, -1 = root
    linktype: np.ndarray                     # (L,) u8 REVOLUTE / PRISMATIC
    axis: np.ndarray                         # (L,3) f64 unit axes (local)
    T0: np.ndarray                           # (L,12) f64 T0_Parent (base transform pre-multiplied into roots)
    qmin: np.ndarray                         # (L,)
    qmax: np.ndarray                         # (L,)
    link_geom: List[int]                     # geometry index per link, -1 = none
    joint_type: Optional[np.ndarray] = None  # (nj,) u8 ; default one Normal joint per link
    joint_link: Optional[np.ndarray] = None  # (nj,) i32
    joint_base: Optional[np.ndarray] = None  # (nj,) i32 RobotModelJoint::baseIndex (-1 = world); default: the parent of joint_link
    drivers: List[DriverSpec] = field(default_factory=list)
    # edits applied after InitAllSelfCollisions (Cpp/Modeling/Robot.cpp:1274-1313): (i, j, enabled)
    self_collision_edits: List[Tuple[int, int, bool]] = field(default_factory=list)
    names: Optional[List[str]] = None

    @property
    def L(self) -> int:
        return int(len(self.parents))


@dataclass
class WorldSpec:
    geoms: List[GeomSpec] = field(default_factory=list)
    terrains: List[int] = field(default_factory=list)               # geometry index per terrain
    objects: List[Tuple[int, np.ndarray]] = field(default_factory=list)  # (geometry index, T12) per rigid object
    robot: Optional[RobotSpec] = None
    pair_mask: Optional[np.ndarray] = None                           # (n_ids,n_ids) u8 overrides InitializeDefault
    world_ids: Optional[np.ndarray] = None                           # engine id -> id in the caller's world (set when other robots' links ride along as rigid objects)

    def add_geom(self, g: GeomSpec) -> int:
        self.geoms.append(g)
        return len(self.geoms) - 1

    def num_ids(self) -> int:
        n = len(self.terrains) + len(self.objects)
        if self.robot is not None:
            n += 1 + self.robot.L
        return n

    # ID helpers (Cpp/Modeling/World.cpp:110-180)
    def terrain_id(self, i: int) -> int:
        return i

    def rigid_object_id(self, i: int) -> int:
        return len(self.terrains) + i

    def robot_id(self) -> int:
        return len(self.terrains) + len(self.objects)

    def robot_link_id(self, j: int) -> int:
        return len(self.terrains) + len(self.objects) + 1 + j

    def total_tris(self) -> int:
        return sum(g.tris.shape[0] for g in self.geoms if g.kind == "mesh")
