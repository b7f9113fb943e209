"""Planning on a subset of the degrees of freedom: the mirror of ``klampt.plan.cspaceutils.EmbeddedCSpace`` (reference
Python/klampt/plan/cspaceutils.py:108-203) -- the Python-side counterpart of ``SingleRobotCSpace::FixDof`` -- with the
two batch entry points the batched planners (klampt_b200.plan.MotionPlan) call.

An embedded configuration holds the values of the DOFs in ``mapping``; ``lift`` writes them into a copy of the ambient configuration
``xinit`` (all other DOFs stay there), ``project`` reads them back.  Every query is answered by the ambient space on the lifted
configurations, so ``feasible_batch`` / ``visible_batch`` on a ``RobotCSpace`` ambient space are still one launch per batch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .cspace import CSpace


class EmbeddedCSpace(CSpace):
    def __init__(self, ambientspace: CSpace, subset: Sequence[int], xinit: Optional[Sequence[float]] = None):
        CSpace.__init__(self)
        self.ambientspace = ambientspace
        n = len(ambientspace.bound)
        self.mapping = list(subset)
        self.xinit = [0.0] * n if xinit is None else list(xinit)           # the zero configuration unless told otherwise
        if len(self.xinit) != n:
            raise ValueError("Invalid length of ambient space vector: %d should be %d" % (len(self.xinit), n))
        self.eps = ambientspace.eps
        self.bound = [ambientspace.bound[i] for i in self.mapping]
        self.properties = ambientspace.properties                          # shared with the ambient space, as in the reference
        if ambientspace.feasibilityTests is not None:
            self.feasibilityTests = [(lambda x, f=f: f(self.lift(x))) for f in ambientspace.feasibilityTests]
            self.feasibilityTestNames = list(ambientspace.feasibilityTestNames)
            self.feasibilityTestDependencies = list(ambientspace.feasibilityTestDependencies)
        if hasattr(ambientspace, "visible"):
            self.visible = lambda a, b: ambientspace.visible(self.lift(a), self.lift(b))

    # ------------------------------------------------------------------ embedding
    def project(self, xamb) -> List[float]:
        if len(xamb) != len(self.xinit):
            raise ValueError("Invalid length of ambient space vector: %d should be %d" % (len(xamb), len(self.xinit)))
        return [xamb[i] for i in self.mapping]

    def lift(self, xemb) -> List[float]:
        if len(xemb) != len(self.mapping):
            raise ValueError("Invalid length of embedded space vector: %d should be %d" % (len(xemb), len(self.mapping)))
        xamb = list(self.xinit)
        for i, j in enumerate(self.mapping):
            xamb[j] = xemb[i]
        return xamb

    def liftPath(self, path):
        return [self.lift(q) for q in path]

    def projectPath(self, path_amb):
        return [self.project(q) for q in path_amb]

    def lift_batch(self, X) -> np.ndarray:
        X = np.asarray(X, dtype=np.float64).reshape(-1, len(self.mapping))
        out = np.tile(np.asarray(self.xinit, dtype=np.float64), (len(X), 1))
        out[:, self.mapping] = X
        return out

    def project_batch(self, Xamb) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(Xamb, dtype=np.float64).reshape(-1, len(self.xinit))[:, self.mapping])

    # ------------------------------------------------------------------ queries, answered by the ambient space
    def feasible(self, x) -> bool:
        return self.ambientspace.feasible(self.lift(x))

    def sample(self):
        return self.project(self.ambientspace.sample())

    def sampleneighborhood(self, c, r):
        return self.project(self.ambientspace.sampleneighborhood(self.lift(c), r))

    def distance(self, a, b) -> float:
        return self.ambientspace.distance(self.lift(a), self.lift(b))

    def interpolate(self, a, b, u):
        return self.project(self.ambientspace.interpolate(self.lift(a), self.lift(b), u))

    def feasible_batch(self, X, **kw):
        return self.ambientspace.feasible_batch(self.lift_batch(X), **kw)

    def visible_batch(self, A, B, **kw):
        return self.ambientspace.visible_batch(self.lift_batch(A), self.lift_batch(B), **kw)


class EmbeddedMotionPlan:
    """MotionPlan over an EmbeddedCSpace that speaks ambient configurations (reference plan/cspaceutils.py:491-561): endpoints and
    milestones are projected on the way in, paths and roadmap vertices lifted on the way out."""

    def __init__(self, space, q0=None, type: Optional[str] = None, **options):
        from .plan import MotionPlan
        if not hasattr(space, "project") or not hasattr(space, "lift"):
            raise ValueError("space argument must have the project and lift methods")
        self.space = space
        self.plan = MotionPlan(space, type, **options)

    def setEndpoints(self, start, goal):
        self.plan.setEndpoints(self.space.project(start), self.space.project(goal))

    def addMilestone(self, x) -> int:
        return self.plan.addMilestone(self.space.project(x))

    def planMore(self, iterations: int):
        self.plan.planMore(iterations)

    def getPath(self, milestone1=None, milestone2=None):
        p = self.plan.getPath(milestone1, milestone2)
        return None if p is None else [self.space.lift(q) for q in p]

    def getSolutionPath(self):
        return self.getPath()

    def getRoadmap(self):
        V, E = self.plan.getRoadmap()
        return [self.space.lift(v) for v in V], E

    def pathCost(self, p) -> float:
        return self.plan.pathCost([self.space.project(x) for x in p])

    def getStats(self) -> dict:
        return self.plan.getStats()

    def close(self):
        self.plan.close()


class AffineEmbeddedCSpace(CSpace):
    """Planning in DRIVER space: ambient configuration q = A x + b (reference plan/cspaceutils.py:205-421).  For robots whose links are
    coupled by affine drivers (URDF mimic joints, grippers) the reference's ``make_space`` plans over the driver values so that coupled
    links move together; ``from_drivers`` builds A and b from the drivers as ``fromRobotDrivers`` (:320-369) does -- a normal driver
    contributes a 1 at (its link, its column), an affine driver its scale per affected link and the offset in b.

    ``project`` is the least-squares inverse, as in the reference.  Bounds: the interval of driver values for which every affected
    link stays inside its ambient bounds, intersected with the driver's own limits when given.  (The reference's "naive bounds"
    (:300-311) drop the offset and store the interval reversed; the interval computed here is the one that code is after.)"""

    def __init__(self, ambientspace: CSpace, A, b=None, driver_limits=None):
        CSpace.__init__(self)
        self.ambientspace = ambientspace
        self.A = np.asarray(A, dtype=np.float64)
        self.n, self.m = self.A.shape
        if self.n != len(ambientspace.bound):
            raise ValueError("Coefficient matrix must have n rows")
        self.b = np.zeros(self.n) if b is None else np.asarray(b, dtype=np.float64)
        if len(self.b) != self.n:
            raise ValueError("Offset matrix must have n entries")
        self.eps = ambientspace.eps
        self.properties = ambientspace.properties
        lo, hi = np.full(self.m, -np.inf), np.full(self.m, np.inf)
        for i, j in zip(*np.nonzero(self.A)):
            v = self.A[i, j]
            e = sorted(((ambientspace.bound[i][0] - self.b[i]) / v, (ambientspace.bound[i][1] - self.b[i]) / v))
            lo[j], hi[j] = max(lo[j], e[0]), min(hi[j], e[1])
        if driver_limits is not None:
            for j, (a, c) in enumerate(driver_limits):
                lo[j], hi[j] = max(lo[j], a), min(hi[j], c)
        self.bound = [(float(a), float(c)) for a, c in zip(lo, hi)]
        if ambientspace.feasibilityTests is not None:
            self.feasibilityTests = [(lambda x, f=f: f(self.lift(x))) for f in ambientspace.feasibilityTests]
            self.feasibilityTestNames = list(ambientspace.feasibilityTestNames)
            self.feasibilityTestDependencies = list(ambientspace.feasibilityTestDependencies)
        if hasattr(ambientspace, "visible"):
            self.visible = lambda a, c: ambientspace.visible(self.lift(a), self.lift(c))

    @staticmethod
    def from_drivers(ambientspace: CSpace, drivers, n_links: int, active: Optional[Sequence[int]] = None) -> "AffineEmbeddedCSpace":
        """`drivers`: DriverSpec-like objects (links, scale, offset, qmin, qmax); `active`: the driver indices to plan over (default all)"""
        idx = list(range(len(drivers))) if active is None else list(active)
        A, b = np.zeros((n_links, len(idx))), np.zeros(n_links)
        for col, d in enumerate(drivers[k] for k in idx):
            for link, s, o in zip(d.links, d.scale, d.offset):
                A[link, col] = s
                b[link] = o
        return AffineEmbeddedCSpace(ambientspace, A, b, driver_limits=[(drivers[k].qmin, drivers[k].qmax) for k in idx])

    fromRobotDrivers = from_drivers

    # ------------------------------------------------------------------ embedding
    def lift(self, xemb) -> List[float]:
        if len(xemb) != self.m:
            raise ValueError("Invalid length of embedded space vector: %d should be %d" % (len(xemb), self.m))
        return list(self.b + self.A @ np.asarray(xemb, dtype=np.float64))

    def project(self, xamb) -> List[float]:
        if len(xamb) != self.n:
            raise ValueError("Invalid length of ambient space vector: %d should be %d" % (len(xamb), self.n))
        return list(np.linalg.lstsq(self.A, np.asarray(xamb, dtype=np.float64) - self.b, rcond=None)[0])

    def liftPath(self, path):
        return [self.lift(q) for q in path]

    def projectPath(self, path_amb):
        return [self.project(q) for q in path_amb]

    def lift_batch(self, X) -> np.ndarray:
        return np.asarray(X, dtype=np.float64).reshape(-1, self.m) @ self.A.T + self.b

    def project_batch(self, Xamb) -> np.ndarray:
        return np.linalg.lstsq(self.A, (np.asarray(Xamb, dtype=np.float64).reshape(-1, self.n) - self.b).T, rcond=None)[0].T

    # ------------------------------------------------------------------ queries, answered by the ambient space
    def feasible(self, x) -> bool:
        return self.ambientspace.feasible(self.lift(x))

    def sample(self):
        return self.project(self.ambientspace.sample())

    def distance(self, a, c) -> float:
        return self.ambientspace.distance(self.lift(a), self.lift(c))

    def interpolate(self, a, c, u):
        return self.project(self.ambientspace.interpolate(self.lift(a), self.lift(c), u))

    def feasible_batch(self, X, **kw):
        return self.ambientspace.feasible_batch(self.lift_batch(X), **kw)

    def visible_batch(self, A, B, **kw):
        return self.ambientspace.visible_batch(self.lift_batch(A), self.lift_batch(B), **kw)
