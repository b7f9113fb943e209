"""Batch-aware planner front end.

The reference's planners call ``CSpace::IsFeasible`` once per sample and ``PathChecker(a,b)->IsVisible()`` once per
candidate edge (Python/klampt/src/motionplanning.cpp:1327-1334); on a GPU engine that leaves the device idle.  This
module keeps the user-facing shape of ``klampt.plan.cspace.MotionPlan`` (reference Python/klampt/plan/cspace.py:216-427:
``MotionPlan(space, type, **options)``, ``setEndpoints``, ``addMilestone``, ``planMore``, ``getPath``, ``getRoadmap``,
``getStats``, ``close``) but runs roadmap planners whose inner loops are the two batch calls:

  'prm'       every round: sample a batch -> feasible_batch -> k nearest neighbours -> visible_batch on all candidate edges
  'prm*'      the same with k = e (1 + 1/d) log n neighbours (asymptotically optimal; an explicit ``knn`` is a lower limit)
  'lazyprm*'  same sampling, but edges are only checked (in batches) when they lie on the current best path
  'rrt'       bidirectional, batch-synchronous RRT: a batch of random targets, each pulls the nearest vertex of the start or the
              goal tree (alternating) one ``perturbationRadius`` step towards it; the new configurations go through
              feasible_batch, their tree edges through visible_batch, and every new vertex then tries to bridge to the nearest
              vertex of the other tree (``connectionThreshold``) with one more visible_batch
  'sbl'       single-query bidirectional lazy planner: both trees grow by feasible samples drawn in the neighbourhood
              (``perturbationRadius``) of random tree vertices WITHOUT checking the tree edges; bridges are proposed between
              close vertices of the two trees; the edges of a candidate start-goal path are validated in one batch when the
              path is asked for, and blocked edges are dropped

Option ``shortcut`` (the reference's ``MotionPlan.setOptions(shortcut=1)``, plan/cspace.py:254-256: "perform shortcutting after
a first plan is found"): once a path exists every further planMore iteration is one batch-synchronous shortcutting round --
``batch`` random pairs of points ON the current path (not only its vertices), all chords checked by one visible_batch, the
visible ones applied greedily by saving as long as their spans do not overlap.

Options (MotionPlan.setOptions keys that apply): ``knn``, ``connectionThreshold``, ``perturbationRadius``, ``shortcut``; plus
``batch`` (samples per planMore iteration) and ``seed``.  The planners need a space with ``feasible_batch(Q)`` and ``visible_batch(A, B)`` (``RobotCSpace``), and use
its ``distance`` metric restricted to the L2 form the engine implements.
"""
from __future__ import annotations

import heapq
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class MotionPlan:
    _next_options: Dict[str, float] = {}

    @staticmethod
    def setOptions(**opts):
        MotionPlan._next_options.update(opts)

    def __init__(self, space, type: Optional[str] = None, **options):
        if not (hasattr(space, "feasible_batch") and hasattr(space, "visible_batch")):
            raise TypeError("MotionPlan needs a space with feasible_batch / visible_batch (klampt_b200.robotcspace.RobotCSpace)")
        opts = dict(MotionPlan._next_options)
        opts.update(options)
        MotionPlan._next_options = {}
        self.space = space
        self.type = (type or "prm").lower()
        if self.type not in ("prm", "prm*", "lazyprm*", "lazyprm", "rrt", "sbl"):
            raise ValueError("planner type %r is not batched here; available: prm, prm*, lazyprm*, rrt, sbl" % type)
        self.lazy = self.type.startswith("lazy") or self.type == "sbl"
        self.tree: List[int] = []                          # rrt / sbl: 0 = vertex of the start tree, 1 = of the goal tree
        self._user_opts = set(opts)
        self.knn = int(opts.get("knn", 10))
        self.connectionThreshold = float(opts.get("connectionThreshold", float("inf")))
        self.batch = int(opts.get("batch", 2048))
        self.perturbationRadius = float(opts.get("perturbationRadius", 0.25))
        self.shortcut = bool(opts.get("shortcut", False))
        self._best: Optional[np.ndarray] = None            # shortcut mode: the current solution, a polyline of configurations
        self.rng = np.random.default_rng(int(opts.get("seed", 0)))
        lo, hi = np.array([b[0] for b in space.bound], dtype=np.float64), np.array([b[1] for b in space.bound], dtype=np.float64)
        # sampling box: an unbounded side is replaced by one turn next to the bounded one (lo + [0, 2 pi) / hi - [0, 2 pi)), or [0, 2 pi)
        self._lo = np.where(np.isfinite(lo), lo, np.where(np.isfinite(hi), hi - 2 * np.pi, 0.0))
        self._hi = np.where(np.isfinite(hi), hi, self._lo + 2 * np.pi)
        # Metric and interpolation.  The engine checks edges along Klampt::Interpolate geodesics (Spin: short arc, Floating / BallAndSocket:
        # SO(3)); for robots with only Normal / Weld joints those are straight lines and the Euclidean forms below are exact and
        # vectorised.  Otherwise every distance / interpolation goes through space.distance / space.interpolate, row by row, so that
        # nearest neighbours, steering, shortcut end points and path costs live on the same curves the edge checker verifies.
        jt = getattr(getattr(getattr(space, "spec", None), "robot", None), "joint_type", None)
        self._geodesic = jt is not None and any(int(t) not in (0, 1) for t in jt)
        self.V = np.zeros((0, len(lo)))                    # milestones
        self.adj: List[Dict[int, Tuple[float, bool]]] = []  # neighbour -> (length, checked)
        self.start = self.goal = None
        self.stats = {"samples": 0, "feasible_samples": 0, "edges_checked": 0, "edges_visible": 0, "iterations": 0,
                      "shortcuts_tried": 0, "shortcuts_applied": 0}

    # ------------------------------------------------------------------ roadmap primitives
    def addMilestone(self, q: Sequence[float]) -> int:
        q = np.asarray(q, dtype=np.float64).reshape(1, -1)
        if not bool(self.space.feasible_batch(q)[0]):
            raise RuntimeError("milestone is infeasible")
        return self._add_vertices(q)[0]

    def _add_vertices(self, Q: np.ndarray) -> List[int]:
        base = len(self.V)
        self.V = np.vstack([self.V, Q])
        self.adj.extend({} for _ in range(len(Q)))
        return list(range(base, base + len(Q)))

    def setEndpoints(self, start: Sequence[float], goal: Sequence[float]):
        ends = np.array([start, goal], dtype=np.float64)
        ok = self.space.feasible_batch(ends)
        if not ok[0]:
            raise RuntimeError("Start configuration is infeasible")
        if not ok[1]:
            raise RuntimeError("Goal configuration is infeasible")
        self.start, self.goal = self._add_vertices(ends)
        if self.type in ("rrt", "sbl"):
            self.tree = [0, 1]
            return
        self._connect([self.start, self.goal])

    def _dist(self, A: np.ndarray, B: np.ndarray) -> np.ndarray:
        if not self._geodesic:
            return np.sqrt(((A - B) ** 2).sum(axis=-1))
        A2, B2 = np.broadcast_arrays(np.atleast_2d(A), np.atleast_2d(B))
        d = np.array([self.space.distance(list(a), list(b)) for a, b in zip(A2, B2)])
        return d if np.ndim(A) > 1 or np.ndim(B) > 1 else d[0]

    def _interp(self, A: np.ndarray, B: np.ndarray, u) -> np.ndarray:
        """Klampt::Interpolate(A[i], B[i], u[i]) per row"""
        u = np.broadcast_to(np.asarray(u, dtype=np.float64).reshape(-1), (len(A),))
        if not self._geodesic:
            return A * (1.0 - u[:, None]) + B * u[:, None]
        return np.array([self.space.interpolate(list(a), list(b), float(t)) for a, b, t in zip(A, B, u)])

    def _candidates(self, new: List[int]) -> np.ndarray:
        """k nearest neighbours of every new vertex among all vertices (brute-force on the host: the roadmap is small next to
        the collision work)"""
        from scipy.spatial import cKDTree
        if len(self.V) < 2:
            return np.zeros((0, 2), dtype=np.int64)
        tree = cKDTree(self.V)
        knn = self.knn
        if self.type.endswith("*"):      # PRM* / Lazy-PRM*: k grows as e (1 + 1/d) log n, the rate that keeps the roadmap asymptotically optimal
            knn = max(1, int(np.ceil(np.e * (1.0 + 1.0 / self.V.shape[1]) * np.log(max(len(self.V), 2)))))
            if "knn" in self._user_opts:
                knn = max(knn, self.knn)
        k = min(knn + 1, len(self.V))
        d, idx = tree.query(self.V[new], k=k)
        d, idx = np.atleast_2d(d), np.atleast_2d(idx)
        pairs = set()
        for row, i in enumerate(new):
            for dist, j in zip(d[row], idx[row]):
                if j != i and np.isfinite(dist) and int(j) not in self.adj[i]:
                    pairs.add((min(i, int(j)), max(i, int(j))))
        cand = np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)
        # The k-d tree ranks by Euclidean distance -- a candidate generator only.  The connection threshold is applied in the space's own
        # metric (identical for Normal-joint robots; for Spin / Floating joints the geodesic can be shorter than the chord in R^n).
        if len(cand) and np.isfinite(self.connectionThreshold):
            cand = cand[self._dist(self.V[cand[:, 0]], self.V[cand[:, 1]]) <= self.connectionThreshold]
        return cand

    def _connect(self, new: List[int]):
        cand = self._candidates(new)
        if len(cand) == 0:
            return
        A, B = self.V[cand[:, 0]], self.V[cand[:, 1]]
        length = self._dist(A, B)
        if self.lazy:
            for (i, j), l in zip(cand, length):
                self.adj[i][int(j)] = (float(l), False)
                self.adj[j][int(i)] = (float(l), False)
            return
        vis = np.asarray(self.space.visible_batch(A, B)).astype(bool)
        self.stats["edges_checked"] += len(cand)
        self.stats["edges_visible"] += int(vis.sum())
        for (i, j), l, v in zip(cand, length, vis):
            if v:
                self.adj[i][int(j)] = (float(l), True)
                self.adj[j][int(i)] = (float(l), True)

    # ------------------------------------------------------------------ planning
    # ------------------------------------------------------------------ tree planners
    def _nearest_in_tree(self, t: int, Q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        from scipy.spatial import cKDTree
        ids = np.nonzero(np.asarray(self.tree) == t)[0]
        d, k = cKDTree(self.V[ids]).query(Q)          # Euclidean ranking picks the tree vertex; the step length below uses the space's metric
        if self._geodesic:
            d = self._dist(self.V[ids[k]], Q)
        return ids[k], d

    def _edge(self, i: int, j: int, checked: bool):
        l = float(self._dist(self.V[i], self.V[j]))
        self.adj[i][j] = (l, checked)
        self.adj[j][i] = (l, checked)

    def _bridge(self, new: List[int]):
        """every new vertex proposes an edge to the nearest vertex of the other tree within connectionThreshold"""
        if not new:
            return
        new_arr = np.asarray(new)
        cand = []
        for t in (0, 1):
            mine = new_arr[np.asarray(self.tree)[new_arr] == t]
            if len(mine) == 0 or not any(x == 1 - t for x in self.tree):
                continue
            other, d = self._nearest_in_tree(1 - t, self.V[mine])
            thr = self.connectionThreshold if np.isfinite(self.connectionThreshold) else (2 * self.perturbationRadius if self.type == "sbl" else np.inf)
            cand += [(int(i), int(j)) for i, j, dd in zip(mine, other, d) if dd <= thr and int(j) not in self.adj[int(i)]]
        if not cand:
            return
        if self.lazy:
            for i, j in cand:
                self._edge(i, j, False)
            return
        A, B = self.V[[i for i, _ in cand]], self.V[[j for _, j in cand]]
        vis = np.asarray(self.space.visible_batch(A, B)).astype(bool)
        self.stats["edges_checked"] += len(cand)
        self.stats["edges_visible"] += int(vis.sum())
        for (i, j), v in zip(cand, vis):
            if v:
                self._edge(i, j, True)

    def _grow_trees(self):
        if self.start is None:
            raise RuntimeError("setEndpoints first")
        n, dim = self.batch, len(self._lo)
        side = (np.arange(n) + self.stats["iterations"]) % 2
        if self.type == "rrt":
            target = self.rng.uniform(self._lo, self._hi, size=(n, dim))
            src = np.empty(n, dtype=np.int64)
            Qn = np.empty((n, dim))
            for t in (0, 1):
                m = side == t
                near, d = self._nearest_in_tree(t, target[m])
                step = np.minimum(1.0, self.perturbationRadius / np.maximum(d, 1e-300))[:, None]
                src[m] = near
                Qn[m] = self._interp(self.V[near], target[m], step[:, 0])
        else:                                   # sbl: a sample in the neighbourhood of a random vertex of the tree
            tree = np.asarray(self.tree)
            src = np.array([self.rng.choice(np.nonzero(tree == t)[0]) for t in side], dtype=np.int64)
            Qn = np.clip(self.V[src] + self.rng.uniform(-self.perturbationRadius, self.perturbationRadius, size=(n, dim)), self._lo, self._hi)
        ok = np.asarray(self.space.feasible_batch(Qn)).astype(bool)
        self.stats["samples"] += n
        self.stats["feasible_samples"] += int(ok.sum())
        keep = ok.copy()
        if self.type == "rrt" and ok.any():      # rrt validates its tree edges now; sbl leaves them for getPath
            vis = np.asarray(self.space.visible_batch(self.V[src[ok]], Qn[ok])).astype(bool)
            self.stats["edges_checked"] += int(ok.sum())
            self.stats["edges_visible"] += int(vis.sum())
            keep[np.nonzero(ok)[0][~vis]] = False
        if not keep.any():
            return
        new = self._add_vertices(Qn[keep])
        self.tree.extend(int(t) for t in side[keep])
        for v, u in zip(new, src[keep]):
            self._edge(v, int(u), self.type == "rrt")
        self._bridge(new)

    def _shortcut_round(self) -> int:
        """one batch of random chords across the current solution; returns how many were applied"""
        P = self._best
        seg = self._dist(P[:-1], P[1:])
        s = np.concatenate([[0.0], np.cumsum(seg)])
        if len(P) < 3 or not s[-1] > 0:
            return 0
        t = np.sort(self.rng.uniform(0.0, s[-1], size=(self.batch, 2)), axis=1)
        ia = np.minimum(np.searchsorted(s, t[:, 0], side="right") - 1, len(seg) - 1)
        ib = np.minimum(np.searchsorted(s, t[:, 1], side="right") - 1, len(seg) - 1)
        def point(i, tt):
            u = np.where(seg[i] > 0, (tt - s[i]) / np.where(seg[i] > 0, seg[i], 1.0), 0.0)
            return self._interp(P[i], P[i + 1], u)
        Xa, Xb = point(ia, t[:, 0]), point(ib, t[:, 1])
        saving = (t[:, 1] - t[:, 0]) - self._dist(Xa, Xb)
        cand = np.nonzero((ib > ia) & (saving > 1e-9 * s[-1]))[0]
        if len(cand) == 0:
            return 0
        vis = np.asarray(self.space.visible_batch(Xa[cand], Xb[cand])).astype(bool)
        self.stats["shortcuts_tried"] += len(cand)
        self.stats["edges_checked"] += len(cand)
        self.stats["edges_visible"] += int(vis.sum())
        good = cand[vis]
        chosen: List[int] = []
        for k in good[np.argsort(-saving[good])]:           # largest saving first, spans must not overlap
            if all(t[k, 1] <= t[c, 0] or t[k, 0] >= t[c, 1] for c in chosen):
                chosen.append(int(k))
        if not chosen:
            return 0
        chosen.sort(key=lambda k: t[k, 0])
        out, nxt = [], 0                                     # nxt: first vertex of P not yet emitted
        for k in chosen:
            out.extend(P[nxt:ia[k] + 1]); out.append(Xa[k]); out.append(Xb[k])
            nxt = ib[k] + 1
        out.extend(P[nxt:])
        Q = np.array(out)
        keep = np.concatenate([[True], self._dist(Q[:-1], Q[1:]) > 0])      # a chord that starts on a vertex repeats it
        self._best = Q[keep]
        self.stats["shortcuts_applied"] += len(chosen)
        return len(chosen)

    def planMore(self, iterations: int):
        if self.shortcut:
            for _ in range(int(iterations)):
                if self._best is None:
                    self._plan_once()
                    path = self._search_path(self.start, self.goal) if self.start is not None else None
                    if path is not None:
                        self._best = np.array(path, dtype=np.float64)
                else:
                    self._shortcut_round()
                    self.stats["iterations"] += 1
            return
        for _ in range(int(iterations)):
            self._plan_once()

    def _plan_once(self):
        if self.type in ("rrt", "sbl"):
            self._grow_trees()
            self.stats["iterations"] += 1
            return
        Q = self.rng.uniform(self._lo, self._hi, size=(self.batch, len(self._lo)))
        ok = np.asarray(self.space.feasible_batch(Q)).astype(bool)
        self.stats["samples"] += len(Q)
        self.stats["feasible_samples"] += int(ok.sum())
        self.stats["iterations"] += 1
        if ok.any():
            self._connect(self._add_vertices(Q[ok]))

    def _shortest(self, src: int, dst: int) -> Optional[List[int]]:
        dist = {src: 0.0}
        prev: Dict[int, int] = {}
        heap = [(0.0, src)]
        while heap:
            d, u = heapq.heappop(heap)
            if u == dst:
                path = [u]
                while path[-1] != src:
                    path.append(prev[path[-1]])
                return path[::-1]
            if d > dist.get(u, float("inf")):
                continue
            for v, (l, _) in self.adj[u].items():
                nd = d + l
                if nd < dist.get(v, float("inf")):
                    dist[v] = nd
                    prev[v] = u
                    heapq.heappush(heap, (nd, v))
        return None

    def getPath(self, milestone1: Optional[int] = None, milestone2: Optional[int] = None) -> Optional[List[List[float]]]:
        src = self.start if milestone1 is None else milestone1
        dst = self.goal if milestone2 is None else milestone2
        if src is None or dst is None:
            raise RuntimeError("setEndpoints (or two milestones) first")
        if self.shortcut and self._best is not None and milestone1 is None and milestone2 is None:
            return [list(q) for q in self._best]
        return self._search_path(src, dst)

    def _search_path(self, src: int, dst: int) -> Optional[List[List[float]]]:
        while True:
            path = self._shortest(src, dst)
            if path is None:
                return None
            if not self.lazy:
                return [list(self.V[i]) for i in path]
            # lazy: validate the unchecked edges of the candidate path in one batch, drop the blocked ones, search again
            todo = [(a, b) for a, b in zip(path[:-1], path[1:]) if not self.adj[a][b][1]]
            if not todo:
                return [list(self.V[i]) for i in path]
            A, B = self.V[[a for a, _ in todo]], self.V[[b for _, b in todo]]
            vis = np.asarray(self.space.visible_batch(A, B)).astype(bool)
            self.stats["edges_checked"] += len(todo)
            self.stats["edges_visible"] += int(vis.sum())
            for (a, b), v in zip(todo, vis):
                if v:
                    l = self.adj[a][b][0]
                    self.adj[a][b] = (l, True)
                    self.adj[b][a] = (l, True)
                else:
                    del self.adj[a][b]
                    del self.adj[b][a]

    def getSolutionPath(self):
        return self.getPath()

    def getRoadmap(self) -> Tuple[List[List[float]], List[Tuple[int, int]]]:
        E = [(i, j) for i, nb in enumerate(self.adj) for j in nb if i < j]
        return [list(v) for v in self.V], E

    def pathCost(self, path) -> float:
        P = np.asarray(path, dtype=np.float64)
        return float(self._dist(P[:-1], P[1:]).sum())

    def getStats(self) -> dict:
        out = dict(self.stats)
        out["milestones"] = len(self.V)
        out["edges"] = sum(len(a) for a in self.adj) // 2
        if self._best is not None:
            out["path_cost"] = self.pathCost(self._best)
        return out

    def close(self):
        self.V = np.zeros((0, self.V.shape[1]))
        self.adj = []
        self._best = None
