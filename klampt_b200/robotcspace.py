"""Mirror of ``klampt.plan.robotcspace.RobotCSpace`` (reference Python/klampt/plan/robotcspace.py:11-130) and of the
C++ ``SingleRobotCSpace`` batch entry points the north star adds (``IsFeasibleBatch`` / ``IsVisibleBatch``).

The reference evaluates, per configuration and through several Python<->C++ crossings, the named tests
"joint limits" -> "setconfig" -> "calcbb" -> "self collision" -> "obj collision i" -> "terrain collision i".
Here one engine call evaluates the conjunction of all of them for a whole batch on the GPU:

    space = RobotCSpace(robot, collider)
    ok  = space.feasible_batch(Q)            # (N,) uint8, SingleRobotCSpace::IsFeasible per row
    vis = space.visible_batch(A, B)          # (N,) uint8, EpsilonEdgeChecker(a,b,eps).IsVisible per row

``feasible(q)`` / ``isVisible(a,b)`` stay available for planners that call one configuration at a time."""
from __future__ import annotations

import math
import random
from typing import List, Optional

import numpy as np

from . import collide
from .cspace import CSpace
from .engine import Engine
from .robotsim import RobotModel


class RobotCSpace(CSpace):
    def __init__(self, robot: RobotModel, collider: Optional[collide.WorldCollider] = None, device: int = 0):
        CSpace.__init__(self)
        self.robot = robot
        self.collider = collider
        self.setBounds(list(zip(*robot.getJointLimits())))
        # self.eps stays CSpace's 1e-3, as in the reference's RobotCSpace (tests/golden/ref_cspace.json); robotplanning.make_space
        # is what sets edgeCheckResolution = 1e-2 there
        self.properties["geodesic"] = 1
        self.joint_limit_failures = [0] * len(self.bound)
        if collider is not None:
            spec = collider.world.to_spec(robot.index, pair_mask=collider.to_pair_mask())
        else:                                            # self-collisions only
            from .worldspec import WorldSpec
            spec = WorldSpec()
            spec.robot = robot.to_spec(spec)
        self.spec = spec
        self.engine = Engine(spec, device=device)
        self._extra: List = []                           # user constraints (addConstraint): host callables, one configuration at a time
        # the named tests of the reference, each answered by the engine for one configuration
        self.addFeasibilityTest(lambda x: self.inJointLimits(x), "joint limits")
        self.addFeasibilityTest(lambda x: bool(self.engine.feasible_batch(np.asarray(x, dtype=np.float64))[0]), "collision free")

    # ------------------------------------------------------------------ batch entry points
    def feasible_batch(self, Q, return_pairs: bool = False):
        """SingleRobotCSpace::IsFeasible per row.  Constraints added with addConstraint are host callables: they run only on the rows
        the engine found feasible (a row they reject keeps first pair (-1, -1))"""
        res = self.engine.feasible_batch(Q, return_pairs=return_pairs)
        ok = res[0] if return_pairs else res
        Q2 = np.asarray(Q, dtype=np.float64).reshape(len(ok), -1)
        # The reference's Python RobotCSpace tests EVERY dimension against self.bound ("joint limits", robotcspace.py:31-75), the engine
        # only Normal / Weld joints and drivers (SingleRobotCSpace::CheckJointLimits).  One contract for isFeasible, feasible and the batch
        # forms: the all-dimension check is ANDed in here (it also honours a setBounds() that tightened the box after construction;
        # the engine keeps enforcing the robot's own limits, so widening the box beyond them has no effect -- as in the C++ space).
        ok &= self._bounds_ok(Q2).astype(ok.dtype)
        if not self._extra:
            return res
        for i in np.nonzero(ok)[0]:
            if not self._extra_ok(Q2[i]):
                ok[i] = 0
        return res

    def _extra_ok(self, q) -> bool:
        x = list(map(float, q))
        return all(c(x) for c in self._extra)

    def _bounds_ok(self, Q2: np.ndarray) -> np.ndarray:
        """vectorised inJointLimits over every dimension of self.bound (closed intervals), without the failure counters"""
        lo = np.array([b[0] for b in self.bound], dtype=np.float64)
        hi = np.array([b[1] for b in self.bound], dtype=np.float64)
        return ((Q2 >= lo) & (Q2 <= hi)).all(axis=1)

    def visible_batch(self, A, B, eps: Optional[float] = None, return_nchecks: bool = False):
        """EpsilonEdgeChecker(a, b, eps).IsVisible per row.  With user constraints, the edges the engine found visible are walked
        again on the host at the same resolution for those constraints alone (bisection order, robot.interpolate)"""
        eps = self.eps if eps is None else eps
        res = self.engine.edges_visible_batch(A, B, eps=eps, return_nchecks=return_nchecks)
        vis = res[0] if return_nchecks else res
        A2, B2 = np.asarray(A, dtype=np.float64).reshape(len(vis), -1), np.asarray(B, dtype=np.float64).reshape(len(vis), -1)
        # all-dimension bounds (see feasible_batch): midpoints between two in-bound endpoints of a box stay in the box, so only edges
        # with an endpoint outside self.bound need their midpoints looked at -- on the host, with the user constraints' walk
        suspect = ~(self._bounds_ok(A2) & self._bounds_ok(B2))
        if not self._extra and not suspect.any():
            return res
        rows = np.nonzero(vis)[0] if self._extra else np.nonzero(vis.astype(bool) & suspect)[0]
        for i in rows:
            a, b = list(A2[i]), list(B2[i])
            length, segs = self.distance(a, b), 1
            while length > eps and vis[i]:
                segs *= 2
                length *= 0.5
                for k in range(1, segs, 2):
                    x = self.interpolate(a, b, float(k) / segs)
                    if (suspect[i] and not self._bounds_ok(np.asarray(x, dtype=np.float64).reshape(1, -1))[0]) or not self._extra_ok(x):
                        vis[i] = 0
                        break
        return res

    def distance_batch(self, Q, upper_bound: float = float("inf"), include_self: bool = False):
        return self.engine.distance_batch(Q, upper_bound=upper_bound, include_self=include_self)

    # ------------------------------------------------------------------ single-configuration face
    def feasible(self, x) -> bool:
        return bool(self.feasible_batch(np.asarray(x, dtype=np.float64).reshape(1, -1))[0])

    def visible(self, a, b) -> bool:
        return bool(self.visible_batch(np.asarray(a, dtype=np.float64).reshape(1, -1), np.asarray(b, dtype=np.float64).reshape(1, -1))[0])

    def addConstraint(self, checker, name=None):
        """an extra feasibility predicate f(q) -> bool (reference robotcspace.py:77-78); honoured by feasible / feasible_batch /
        visible / visible_batch and listed among the named tests"""
        self.addFeasibilityTest(checker, name)
        self._extra.append(checker)

    def sample(self):
        res = CSpace.sample(self)
        for i, x in enumerate(res):
            if math.isnan(x) or math.isinf(x):
                res[i] = random.uniform(0, math.pi * 2.0)
        return res

    def inJointLimits(self, x) -> bool:
        for i, (xi, bi) in enumerate(zip(x, self.bound)):
            if xi < bi[0] or xi > bi[1]:
                self.joint_limit_failures[i] += 1
                return False
        return True

    def selfCollision(self, x=None) -> bool:
        if x is not None:
            self.robot.setConfig(x)
        return self.robot.selfCollides()

    def envCollision(self, x=None) -> bool:
        if self.collider is None:
            return False
        q = np.asarray(self.robot.getConfig() if x is None else x, dtype=np.float64)
        return self._env_hit(q)

    def _env_hit(self, q) -> bool:
        ok, pair = self.engine.feasible_batch(q, return_pairs=True)
        if ok[0]:
            return False
        rid = self.spec.robot_id()
        if pair[0, 0] < 0:
            return False                                  # infeasible by limits only
        return not (pair[0, 0] > rid and pair[0, 1] > rid) or self._env_only_hit(q)

    def _env_only_hit(self, q) -> bool:
        # the reported pair was a self pair: ask for the environment clearance explicitly
        d = self.engine.distance_batch(q, upper_bound=1e-12, include_self=False)
        return bool(d[0] <= 0.0)

    def colliding_pairs_batch(self, Q, max_pairs: int = 8):
        """all colliding world-id pairs per configuration (no early exit); see Engine.colliding_pairs_batch"""
        return self.engine.colliding_pairs_batch(Q, max_pairs=max_pairs)

    def _counts(self):
        """(terrains, rigid objects of the world, all engine objects): the engine's objects beyond the world's own are the links of
        other robots, which ride along as rigid bodies (WorldModel.to_spec)"""
        T = len(self.spec.terrains)
        O = self.collider.world.numRigidObjects() if self.collider is not None else len(self.spec.objects)
        return T, O, len(self.spec.objects)

    def _other_robot_of(self, k: int):
        """engine object k >= O belongs to another robot: (robot index, name)"""
        w = self.collider.world
        wid = int(self.spec.world_ids[len(self.spec.terrains) + k])
        for r in range(w.numRobots()):
            if w.robotID(r) < wid <= w.robotID(r) + w.robot(r).numLinks():
                return r, w.robot(r).getName()
        return -1, ""

    def feasibilityFailures(self, x):
        """names of the reference's feasibility tests that fail at x (CSpaceInterface::feasibilityFailures with the test names of
        plan/robotcspace.py:62-75): 'joint limits', 'self collision', 'obj collision i name', 'terrain collision i name'"""
        if not self.inJointLimits(x):
            return ["joint limits"]
        pairs, count = self.engine.colliding_pairs_batch(np.asarray(x, dtype=np.float64), max_pairs=32)
        names = []
        T, O, Oall = self._counts()
        for a, b in pairs[0]:
            if a < 0:
                continue
            lo = min(int(a), int(b))
            if lo < T:
                n = "terrain collision %d %s" % (lo, self.collider.world.terrain(lo).getName() if self.collider else "")
            elif lo < T + O:
                n = "obj collision %d %s" % (lo - T, self.collider.world.rigidObject(lo - T).getName() if self.collider else "")
            elif lo < T + Oall:
                n = "robot collision %d %s" % self._other_robot_of(lo - T)
            else:
                n = "self collision"
            if n not in names:
                names.append(n)
        return names

    def feasibilityTestNamesList(self) -> List[str]:
        """the reference's test list for this world, in its order (plan/robotcspace.py:31-75; tests/golden/ref_cspace.json is what the
        reference's own constructor produces): joint limits, the two bookkeeping tests setconfig / calcbb (always true: the engine
        does FK and its own broad phase per configuration), self collision, one test per rigid object and per terrain"""
        if self.collider is None:
            return ["joint limits", "setconfig", "self collision"]
        w = self.collider.world
        names = ["joint limits", "setconfig", "calcbb", "self collision"]
        names += ["obj collision %d %s" % (i, w.rigidObject(i).getName()) for i in range(w.numRigidObjects())]
        names += ["terrain collision %d %s" % (i, w.terrain(i).getName()) for i in range(w.numTerrains())]
        # not in the reference's Python list (its RobotCSpace never looks at other robots); the C++ space does (RobotCSpace.cpp:806-809)
        names += ["robot collision %d %s" % (r, w.robot(r).getName()) for r in range(w.numRobots()) if w.robot(r) is not self.robot]
        return names

    def feasibilityTestDependenciesList(self) -> List[tuple]:
        """(test, prerequisite) pairs as the reference declares them (robotcspace.py:62-72)"""
        if self.collider is None:
            return [("self collision", "setconfig")]
        deps = [("calcbb", "setconfig"), ("self collision", "setconfig")]
        deps += [(n, "calcbb") for n in self.feasibilityTestNamesList() if n.startswith(("obj collision", "terrain collision", "robot collision"))]
        return deps

    def testFeasibility(self, name: str, x) -> bool:
        """one of the reference's named tests at x (CSpaceInterface::testFeasibility): evaluated from the all-pairs query, so any
        number of names costs one launch"""
        if name not in self.feasibilityTestNamesList():
            raise ValueError("Invalid feasibility test name %r" % name)
        if name == "joint limits":
            return self.inJointLimits(x)
        if name in ("setconfig", "calcbb"):
            return True
        pairs, count = self.engine.colliding_pairs_batch(np.asarray(x, dtype=np.float64), max_pairs=32)
        T, O, Oall = self._counts()
        for a, b in pairs[0]:
            if a < 0:
                continue
            lo = min(int(a), int(b))
            if name == "self collision" and lo >= T + Oall:
                return False
            if name.startswith("robot collision ") and T + O <= lo < T + Oall and name == "robot collision %d %s" % self._other_robot_of(lo - T):
                return False
            if name.startswith("terrain collision %d " % lo) and lo < T:
                return False
            if name.startswith("obj collision %d " % (lo - T)) and T <= lo < T + O:
                return False
        return True

    def interpolate(self, a, b, u):
        return self.robot.interpolate(a, b, u)

    def distance(self, a, b):
        return self.robot.distance(a, b)

    def executablePath(self, path):
        return path

    def getStats(self) -> dict:
        out = CSpace.getStats(self)
        out.update({"engine_" + k: v for k, v in self.engine.stats().items()})
        return out
