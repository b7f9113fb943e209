"""Mirror of ``klampt.plan.cspace.CSpace`` (reference Python/klampt/plan/cspace.py:76-214) for the members the
feasibility / visibility path uses: ``eps``, ``bound``, ``properties``, ``addFeasibilityTest(func,name,dependencies)``,
``feasible``, ``isFeasible``, ``isVisible``, ``getStats``, ``setup``, ``close``.  The reference forwards
``isFeasible`` / ``isVisible`` to the C++ ``CSpaceInterface`` (per-test timing, bisection edge checker,
Python/klampt/src/motionplanning.cpp:163-175,383-401,1031-1055); here the same bookkeeping is kept on the host and
subclasses (``RobotCSpace``) route the actual tests to the GPU engine, one configuration or a batch at a time."""
from __future__ import annotations

import random
import time
from typing import Callable, List, Optional, Sequence, Tuple, Union


class CSpace:
    def __init__(self):
        self.cspace = None
        self.feasibilityTests: Optional[List[Callable]] = None
        self.feasibilityTestNames: Optional[List[str]] = None
        self.feasibilityTestDependencies: Optional[List[Tuple[str, str]]] = None
        self.eps = 1e-3
        self.bound: List[Tuple[float, float]] = [(0, 1)]
        self.properties = {}
        self._stats = {"feasible_count": 0, "feasible_true": 0, "feasible_time": 0.0,
                       "visible_count": 0, "visible_true": 0, "visible_time": 0.0, "visible_length": 0.0}
        self._test_stats = {}
        self._adaptive = False
        self._prior = {"feasible": {}, "visible": {}}       # name -> [cost, probability, count] (AdaptiveCSpace::PredicateStats)
        self._order = {"feasible": None, "visible": None}   # optimised test orders (lists of names) or None = as added
        self._vis_deps: List[Tuple[str, str]] = []

    # ------------------------------------------------------------------ bounds / sampling
    def setBounds(self, bound: Sequence[Tuple[float, float]]):
        self.bound = list(bound)
        self.properties["minimum"] = [b[0] for b in bound]
        self.properties["maximum"] = [b[1] for b in bound]
        volume = 1
        for b in self.bound:
            if b[0] != b[1]:
                volume *= b[1] - b[0]
        self.properties["volume"] = volume

    def sample(self):
        return [random.uniform(*b) for b in self.bound]

    def sampleneighborhood(self, c, r):
        return [random.uniform(max(b[0], ci - r), min(b[1], ci + r)) for ci, b in zip(c, self.bound)]

    def inBounds(self, x) -> bool:
        return all(a <= xi <= b for (xi, (a, b)) in zip(x, self.bound))

    # ------------------------------------------------------------------ tests
    def addFeasibilityTest(self, func: Callable, name: Optional[str] = None, dependencies: Optional[Union[str, List[str]]] = None):
        if self.feasibilityTests is None:
            self.feasibilityTests, self.feasibilityTestNames, self.feasibilityTestDependencies = [], [], []
        assert name is None or isinstance(name, str), "Name argument 'name' must be a string"
        assert callable(func), "Feasibility test 'func' must be a callable object"
        self.feasibilityTests.append(func)
        self._order = {"feasible": None, "visible": None}     # an optimised order does not know the new test
        if name is None:
            name = "test_" + str(len(self.feasibilityTests) - 1)
        self.feasibilityTestNames.append(name)
        if dependencies is not None:
            for d in (dependencies if isinstance(dependencies, (list, tuple)) else [dependencies]):
                self.feasibilityTestDependencies.append((name, d))

    def feasible(self, x) -> bool:
        if self.feasibilityTests is None:
            return self.inBounds(x)
        for test in self.feasibilityTests:
            if not test(x):
                return False
        return True

    def feasibilityFailures(self, x) -> List[str]:
        """names of the tests that fail at x (CSpaceInterface::feasibilityFailures)"""
        if self.feasibilityTests is None:
            return [] if self.inBounds(x) else ["bounds"]
        return [n for n, t in zip(self.feasibilityTestNames, self.feasibilityTests) if not t(x)]

    def testFeasibility(self, name: str, x) -> bool:
        """one named test (CSpaceInterface::testFeasibility, Python/klampt/src/motionplanning.h:124)"""
        if self.feasibilityTests is None or name not in self.feasibilityTestNames:
            raise ValueError("Invalid feasibility test name %r" % name)
        return bool(self.feasibilityTests[self.feasibilityTestNames.index(name)](x))

    def feasibilityQueryOrder(self) -> List[str]:
        """the order in which isFeasible runs the named tests (CSpaceInterface::feasibilityQueryOrder): the order they were added
        in until optimizeQueryOrder has run"""
        return list(self._order["feasible"] or self.feasibilityTestNames or [])

    def visibilityQueryOrder(self) -> List[str]:
        return list(self._order["visible"] or self.feasibilityTestNames or [])

    # ------------------------------------------------------------------ adaptive queries (motionplanning.h:139-166)
    # The reference keeps, per named test, a running (cost, success probability, evidence count) and can re-order the tests of the
    # conjunction to minimise its expected cost.  The re-ordering itself is KrisLibrary's AdaptiveCSpace::OptimizeQueryOrder, which
    # is not in the reference tree; the rule used here is the classical one for a conjunction of independent tests -- run them by
    # increasing cost / (1 - p), p = probability of passing -- applied greedily among the tests whose prerequisites are placed.
    def adaptiveQueriesEnabled(self) -> bool:
        return self._adaptive

    def enableAdaptiveQueries(self, enabled: bool = True):
        self._adaptive = self._adaptive or bool(enabled)     # motionplanning.cpp:910-916 never switches it off again

    def _need_adaptive(self):
        if not self._adaptive:
            raise RuntimeError("adaptive queries not enabled for this space")

    def _pstats(self, kind: str, name: str) -> List[float]:
        if self.feasibilityTestNames is None or name not in self.feasibilityTestNames:
            raise ValueError("Invalid constraint name")
        return self._prior[kind].setdefault(name, [0.0, 0.5, 0.0])      # PyCSpace's initial stats: cost 0, probability 0.5, count 0

    @staticmethod
    def _update_stats(s: List[float], cost: float, passed: bool, strength: float = 1.0):
        n = s[2] + strength
        s[0] += (cost - s[0]) * strength / n
        s[1] += (float(passed) - s[1]) * strength / n
        s[2] = n

    def setFeasibilityDependency(self, name: str, precedingTest: str):
        self._need_adaptive()
        if name not in (self.feasibilityTestNames or []) or precedingTest not in (self.feasibilityTestNames or []) or name == precedingTest:
            raise ValueError("Invalid dependency")
        self.feasibilityTestDependencies.append((name, precedingTest))

    def setVisibilityDependency(self, name: str, precedingTest: str):
        self._need_adaptive()
        if name not in (self.feasibilityTestNames or []) or precedingTest not in (self.feasibilityTestNames or []) or name == precedingTest:
            raise ValueError("Invalid dependency")
        self._vis_deps.append((name, precedingTest))

    def setFeasibilityPrior(self, name: str, costPrior: float = 0.0, feasibilityProbability: float = 0.0, evidenceStrength: float = 1.0):
        self._need_adaptive()
        self._pstats("feasible", name)[:] = [float(costPrior), float(feasibilityProbability), float(evidenceStrength)]

    def setVisibilityPrior(self, name: str, costPrior: float = 0.0, visibilityProbability: float = 0.0, evidenceStrength: float = 1.0):
        self._need_adaptive()
        self._pstats("visible", name)[:] = [float(costPrior), float(visibilityProbability), float(evidenceStrength)]

    def feasibilityCost(self, name: str) -> float:
        self._need_adaptive()
        return self._pstats("feasible", name)[0]

    def feasibilityProbability(self, name: str) -> float:
        self._need_adaptive()
        return self._pstats("feasible", name)[1]

    def visibilityCost(self, name: str) -> float:
        self._need_adaptive()
        return self._pstats("visible", name)[0]

    def visibilityProbability(self, name: str) -> float:
        self._need_adaptive()
        return self._pstats("visible", name)[1]

    def optimizeQueryOrder(self):
        self._need_adaptive()
        names = list(self.feasibilityTestNames or [])
        for kind, deps in (("feasible", self.feasibilityTestDependencies or []), ("visible", self._vis_deps)):
            placed, order = set(), []
            def key(n):
                c, p, _ = self._pstats(kind, n)
                return (c / (1.0 - p) if p < 1.0 else float("inf"), names.index(n))
            while len(order) < len(names):
                ready = [n for n in names if n not in placed and all(d in placed for m, d in deps if m == n)]
                if not ready:
                    raise ValueError("Invalid dependency")            # a cycle
                best = min(ready, key=key)
                order.append(best)
                placed.add(best)
            self._order[kind] = order

    def feasibilityTestDependenciesOf(self, name: str) -> List[str]:
        return [d for n, d in (self.feasibilityTestDependencies or []) if n == name]

    def testVisibility(self, name: str, a, b) -> bool:
        """visibility of the straight line under ONE named feasibility test (CSpaceInterface::testVisibility): the epsilon edge checker
        restricted to that test"""
        if self.feasibilityTests is None or name not in self.feasibilityTestNames:
            raise ValueError("Invalid visibility test name %r" % name)
        test = self.feasibilityTests[self.feasibilityTestNames.index(name)]
        length, segs = self.distance(a, b), 1
        while length > self.eps:
            segs *= 2
            length *= 0.5
            for k in range(1, segs, 2):
                if not test(self.interpolate(a, b, float(k) / segs)):
                    return False
        return True

    def visibilityFailures(self, a, b) -> List[str]:
        """names of the tests under which the line from a to b is not visible (CSpaceInterface::visibilityFailures)"""
        if self.feasibilityTests is None:
            return [] if self.isVisible(a, b) else ["visible"]
        return [n for n in self.feasibilityTestNames if not self.testVisibility(n, a, b)]

    def setVisibilityEpsilon(self, eps: float):
        if not eps > 0:
            raise ValueError("Invalid epsilon")          # motionplanning.cpp: PyException("Invalid epsilon")
        self.eps = float(eps)

    def distance(self, a, b) -> float:
        return sum((p - q) ** 2 for p, q in zip(a, b)) ** 0.5

    def interpolate(self, a, b, u):
        return [p * (1.0 - u) + q * u for p, q in zip(a, b)]

    # ------------------------------------------------------------------ the CSpaceInterface face
    def setup(self, reinit: bool = False):
        """the reference builds its CSpaceInterface here and, when the tests are named, enables adaptive queries and registers the
        dependencies (plan/cspace.py:112-118); the host-side bookkeeping of this class plays that part"""
        self.cspace = self
        if self.feasibilityTests is not None:
            self.enableAdaptiveQueries()

    def close(self):
        self.cspace = None

    def isFeasible(self, x) -> bool:
        t0 = time.perf_counter()
        if self.feasibilityTests is None:
            ok = self.feasible(x)
        else:
            ok = True
            order = self._order["feasible"]
            seq = zip(self.feasibilityTestNames, self.feasibilityTests) if order is None else \
                ((n, self.feasibilityTests[self.feasibilityTestNames.index(n)]) for n in order)
            for n, test in seq:
                t1 = time.perf_counter()
                r = bool(test(x))
                dt = time.perf_counter() - t1
                s = self._test_stats.setdefault(n, [0, 0, 0.0])
                s[0] += 1
                s[1] += int(r)
                s[2] += dt
                if self._adaptive:
                    self._update_stats(self._pstats("feasible", n), dt, r)
                if not r:
                    ok = False
                    break
        self._stats["feasible_count"] += 1
        self._stats["feasible_true"] += int(ok)
        self._stats["feasible_time"] += time.perf_counter() - t0
        return ok

    def isVisible(self, a, b) -> bool:
        """EpsilonEdgeChecker on this space's own distance / interpolate / feasible: bisect until the segment length is
        <= eps, midpoints coarse to fine, endpoints not re-checked (reference SURVEY 3.2; resolution ``self.eps``)."""
        t0 = time.perf_counter()
        if hasattr(self, "visible"):
            ok = bool(self.visible(a, b))
        else:
            ok = True
            length = self.distance(a, b)
            segs = 1
            while length > self.eps and ok:
                segs *= 2
                length *= 0.5
                for k in range(1, segs, 2):
                    if not self.feasible(self.interpolate(a, b, float(k) / segs)):
                        ok = False
                        break
        self._stats["visible_count"] += 1
        self._stats["visible_true"] += int(ok)
        self._stats["visible_time"] += time.perf_counter() - t0
        self._stats["visible_length"] += self.distance(a, b)
        return ok

    def getStats(self) -> dict:
        """same keys as CSpaceInterface::getStats (motionplanning.cpp:1031-1055); empty before setup(), as in the reference
        (plan/cspace.py:195-199)"""
        if self.cspace is None:
            return {}
        s = self._stats
        out = {"feasible_count": s["feasible_count"],
               "feasible_probability": s["feasible_true"] / s["feasible_count"] if s["feasible_count"] else 0.0,
               "feasible_time": s["feasible_time"] / s["feasible_count"] if s["feasible_count"] else 0.0,
               "visible_count": s["visible_count"],
               "visible_probability": s["visible_true"] / s["visible_count"] if s["visible_count"] else 0.0,
               "visible_time": s["visible_time"] / s["visible_count"] if s["visible_count"] else 0.0,
               "average_visible_length": s["visible_length"] / s["visible_count"] if s["visible_count"] else 0.0}
        for n, (cnt, tr, tm) in self._test_stats.items():
            out[n + "_count"] = cnt
            out[n + "_probability"] = tr / cnt if cnt else 0.0
            out[n + "_time"] = tm / cnt if cnt else 0.0
        return out
