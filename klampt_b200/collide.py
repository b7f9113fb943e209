"""Mirror of ``klampt.model.collide.WorldCollider`` (reference Python/klampt/model/collide.py:246-362,531-698): the
mask of geometry pairs that planning code wants checked, plus the group iterators.  The mask semantics restated:
terrain-object; terrain-link only if the link has a parent (fixed base links rest on the terrain); object-object;
object-link; links of different robots; self pairs from ``selfCollisionEnabled``; ``ignoreCollision`` edits.

``to_pair_mask()`` turns the mask into the dense ``collisionEnabled`` matrix over world IDs that the engine's
``kb_set_pair_mask`` takes, so ignored pairs are honoured by the batched kernels.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Tuple

import numpy as np

from .robotsim import WorldModel, RobotModel, RobotModelLink, RigidObjectModel, TerrainModel


def bb_intersect(a, b) -> bool:
    """axis-aligned boxes (bmin,bmax) overlap, closed intervals"""
    amin, amax = a
    bmin, bmax = b
    return not any(q < u or v < p for (p, q, u, v) in zip(amin, amax, bmin, bmax))


def bb_create(*ptlist):
    """box of a set of points; no points: the empty box (+inf, -inf)"""
    if not ptlist:
        return [float("inf")] * 3, [float("-inf")] * 3
    return [min(c) for c in zip(*ptlist)], [max(c) for c in zip(*ptlist)]


def bb_empty(bb) -> bool:
    return any(a > b for a, b in zip(bb[0], bb[1]))


def bb_contains(bb, x) -> bool:
    return all(p <= v <= q for p, q, v in zip(bb[0], bb[1], x))


def bb_intersection(*bbs):
    """may be empty (bb_empty)"""
    return [max(x) for x in zip(*[b[0] for b in bbs])], [min(x) for x in zip(*[b[1] for b in bbs])]


def bb_union(*bbs):
    return [min(x) for x in zip(*[b[0] for b in bbs])], [max(x) for x in zip(*[b[1] for b in bbs])]


def _wanted(pairs, i, j) -> bool:
    return pairs == "all" or pairs(i, j)


def self_collision_iter(geomlist, pairs="all") -> Iterator[Tuple[int, int]]:
    """colliding pairs (i, j), i < j, within one list of geometries (reference collide.py:61-105: no box pre-reject here).
    `pairs`: 'all', a predicate f(i, j), or an explicit list of index pairs"""
    if pairs == "all" or callable(pairs):
        cand = ((i, j) for i in range(len(geomlist)) for j in range(i + 1, len(geomlist)) if _wanted(pairs, i, j))
    else:
        cand = iter(pairs)
    for i, j in cand:
        if geomlist[i].collides(geomlist[j]):
            yield (i, j)


def _group_pairs(geoms1, geoms2, ids1, ids2, pairs):
    """the broad phase the group iterators share (collide.py:131-136): a geometry takes part only if its box touches the union box
    of the other group"""
    bb1 = {i: geoms1[i].getBB() for i in ids1}
    bb2 = {j: geoms2[j].getBB() for j in ids2}
    u1, u2 = bb_union(*bb1.values()), bb_union(*bb2.values())
    keep1 = [i for i in ids1 if bb_intersect(bb1[i], u2)]
    keep2 = [j for j in ids2 if bb_intersect(bb2[j], u1)]
    return ((i, j) for i in keep1 for j in keep2 if _wanted(pairs, i, j))


def group_collision_iter(geomlist1, geomlist2, pairs="all") -> Iterator[Tuple[int, int]]:
    """colliding pairs (i, j) between two lists (reference collide.py:107-157); an explicit pair list skips the box pre-reject"""
    if len(geomlist1) == 0 or len(geomlist2) == 0:
        return
    if pairs == "all" or callable(pairs):
        cand = _group_pairs(geomlist1, geomlist2, range(len(geomlist1)), range(len(geomlist2)), pairs)
    else:
        cand = iter(pairs)
    for i, j in cand:
        if geomlist1[i].collides(geomlist2[j]):
            yield (i, j)


def group_subset_collision_iter(geomlist, alist, blist, pairs="all") -> Iterator[Tuple[int, int]]:
    """colliding pairs between two index subsets of one list (reference collide.py:159-216).  The reference only fills the boxes of
    `blist` entries that are also in `alist` (`if bblist[id] is not None`, :191) and fails on the others; here every listed
    geometry gets its box, which is what the docstring there describes."""
    if pairs != "all" and not callable(pairs):
        cand = iter(pairs)
    elif len(alist) == 0 or len(blist) == 0:
        return
    else:
        cand = _group_pairs(geomlist, geomlist, list(alist), list(blist), pairs)
    for i, j in cand:
        if geomlist[i].collides(geomlist[j]):
            yield (i, j)


def ray_cast(geomlist, s, d):
    """first hit of the ray (s, d) with a list of geometries: (index, point) or None (reference collide.py:219-243: the nearest by
    dot(d, pt - s), a strict '<' so that a tie keeps the earlier geometry)"""
    res = None
    dmin = 1e300
    for i, g in enumerate(geomlist):
        (coll, pt) = g.rayCast(s, d)
        if coll:
            dist = float(np.dot(d, np.subtract(pt, s)))
            if dist < dmin:
                dmin, res = dist, (i, pt)
    return res


class WorldCollider:
    def __init__(self, world: WorldModel, ignore=()):
        self.world = world
        self.geomList: List[Tuple[object, object]] = []
        self.mask: List[set] = []
        self.terrains: List[int] = []
        self.rigidObjects: List[int] = []
        self.robots: List[List[int]] = []
        self._ids: List[int] = []                       # world id of each geomList entry

        def add(obj, wid) -> int:
            g = obj.geometry()
            if g is None or g.type() == "":
                return -1
            self.geomList.append((obj, g))
            self._ids.append(wid)
            return len(self.geomList) - 1

        for i in range(world.numTerrains()):
            self.terrains.append(add(world.terrain(i), world.terrainID(i)))
        for i in range(world.numRigidObjects()):
            self.rigidObjects.append(add(world.rigidObject(i), world.rigidObjectID(i)))
        for r in range(world.numRobots()):
            rob = world.robot(r)
            self.robots.append([add(rob.link(j), world.robotLinkID(r, j)) for j in range(rob.numLinks())])
        self.mask = [set() for _ in self.geomList]

        def on(a, b):
            if a >= 0 and b >= 0:
                self.mask[a].add(b)
                self.mask[b].add(a)

        for t in self.terrains:
            for o in self.rigidObjects:
                on(t, o)
            for links in self.robots:
                for l in links:
                    if l >= 0 and self.geomList[l][0].getParent() >= 0:    # fixed links are allowed to touch the terrain
                        on(t, l)
        for o in self.rigidObjects:
            if o < 0:
                continue
            # sic: the reference slices by the geomList INDEX o, not by the object's position (collide.py:334-339), so with terrains in
            # front of the list an object is also paired with itself and with some later objects (the mask is symmetric, and every
            # consumer walks i < j, so only the (o, o) entries are visible -- tests/test_reference_golden.py holds the mirror to it)
            for o2 in self.rigidObjects[:o]:
                on(o, o2)
            for links in self.robots:
                for l in links:
                    on(o, l)
        for r, links in enumerate(self.robots):
            for other in self.robots[:r]:
                for l1 in links:
                    for l2 in other:
                        on(l1, l2)
            rob = world.robot(r)
            for i in range(rob.numLinks()):
                for j in range(i):
                    if rob.selfCollisionEnabled(i, j):
                        on(links[i], links[j])
        for item in ignore:
            self.ignoreCollision(item)

    # ------------------------------------------------------------------ mask edits
    def _getGeomIndex(self, obj) -> int:
        if isinstance(obj, int):
            return obj
        for i, (o, g) in enumerate(self.geomList):
            if o is obj:
                return i
        return -1

    def ignoreCollision(self, ign):
        """ign: an object (all its pairs are dropped) or a pair of objects"""
        if isinstance(ign, (tuple, list)) and len(ign) == 2:
            a, b = self._getGeomIndex(ign[0]), self._getGeomIndex(ign[1])
            if a < 0 or b < 0:
                raise ValueError("Invalid ignore collision item, must be a pair of bodies in the world")
            self.mask[a].discard(b)
            self.mask[b].discard(a)
        else:
            a = self._getGeomIndex(ign)
            if a < 0:
                raise ValueError("Invalid ignore collision item, must be a body in the world")
            for b in list(self.mask[a]):
                self.mask[b].discard(a)
            self.mask[a] = set()

    def isCollisionEnabled(self, obj_or_pair) -> bool:
        if isinstance(obj_or_pair, (tuple, list)) and len(obj_or_pair) == 2:
            a, b = self._getGeomIndex(obj_or_pair[0]), self._getGeomIndex(obj_or_pair[1])
            return a >= 0 and b >= 0 and b in self.mask[a]
        a = self._getGeomIndex(obj_or_pair)
        return a >= 0 and len(self.mask[a]) > 0

    def to_pair_mask(self) -> np.ndarray:
        """dense collisionEnabled matrix over world IDs (row-major, symmetric) for kb_set_pair_mask"""
        n = self.world.numIDs()
        m = np.zeros((n, n), dtype=np.uint8)
        for a, s in enumerate(self.mask):
            for b in s:
                ia, ib = self._ids[a], self._ids[b]
                m[min(ia, ib), max(ia, ib)] = 1          # upper triangular like selfCollisions(j,k), j<k ...
                if not (self._is_link(a) and self._is_link(b)):
                    m[max(ia, ib), min(ia, ib)] = 1      # ... and symmetric for robot-vs-environment entries
        return m

    def _is_link(self, a) -> bool:
        return isinstance(self.geomList[a][0], RobotModelLink)

    # ------------------------------------------------------------------ iterators (pairs that collide right now)
    def _colliding(self, a: int, b: int) -> bool:
        return self.geomList[a][1].collides(self.geomList[b][1])

    def collisionTests(self, filter1=None, filter2=None, bb_reject=True) -> Iterator[Tuple[tuple, tuple]]:
        """((object, geometry), (object, geometry)) pairs that should be tested, as the reference enumerates them (collide.py:429-499):
        no filter -- every enabled pair once, lower geomList index first, boxes pre-rejected if `bb_reject`; filter1 only -- pairs
        within the set filter1 accepts (no box pre-reject there: a TODO in the reference); both -- pairs between the two sets,
        oriented (member of set 1, member of set 2).  Two deliberate differences: a body is never paired with itself (the
        reference's mask holds (o, o) entries for rigid objects, see __init__, and its first branch lets them through), and with two
        filters a pair is listed once (the reference lists it from both sides)."""
        if filter1 is None:
            bbs = [g[1].getBB() for g in self.geomList] if bb_reject else None
            for a, s in enumerate(self.mask):
                for b in s:
                    if b > a and not (bb_reject and not bb_intersect(bbs[a], bbs[b])):
                        yield self.geomList[a], self.geomList[b]
        elif filter2 is None:
            for a, s in enumerate(self.mask):
                if filter1(self.geomList[a][0]):
                    for b in s:
                        if b > a and filter1(self.geomList[b][0]):
                            yield self.geomList[a], self.geomList[b]
        else:
            for a, s in enumerate(self.mask):
                A = self.geomList[a]
                for b in s:
                    if b > a:
                        B = self.geomList[b]
                        if filter1(A[0]) and filter2(B[0]):
                            yield A, B
                        elif filter1(B[0]) and filter2(A[0]):
                            yield B, A

    def collisions(self, filter1=None, filter2=None):
        for A, B in self.collisionTests(filter1, filter2):
            if A[1].collides(B[1]):
                yield A[0], B[0]

    def robotSelfCollisions(self, robot=None):
        """colliding (RobotModelLink, RobotModelLink) pairs, the higher link first as in the reference (collide.py:531-560); robot =
        None: every robot.  The box pre-reject is this mirror's (it cannot change the answer: getBB contains the geometry)."""
        if robot is None:
            for r in range(len(self.robots)):
                yield from self.robotSelfCollisions(r)
            return
        if isinstance(robot, RobotModel):
            robot = robot.index
        links = self.robots[robot]
        for i, a in enumerate(links):
            for b in links[:i]:
                if a >= 0 and b >= 0 and b in self.mask[a] and bb_intersect(self.geomList[a][1].getBB(), self.geomList[b][1].getBB()) \
                        and self._colliding(a, b):
                    yield self.geomList[a][0], self.geomList[b][0]

    def robotObjectCollisions(self, robot, object=None):
        if isinstance(robot, RobotModel):
            robot = robot.index
        objs = range(len(self.rigidObjects)) if object is None else [object.index if isinstance(object, RigidObjectModel) else object]
        for o in objs:
            b = self.rigidObjects[o]
            if b < 0:
                continue
            for a in self.robots[robot]:
                if a >= 0 and b in self.mask[a] and bb_intersect(self.geomList[a][1].getBB(), self.geomList[b][1].getBB()) and self._colliding(a, b):
                    yield self.geomList[a][0], self.geomList[b][0]

    def robotTerrainCollisions(self, robot, terrain=None):
        if isinstance(robot, RobotModel):
            robot = robot.index
        ters = range(len(self.terrains)) if terrain is None else [terrain.index if isinstance(terrain, TerrainModel) else terrain]
        for t in ters:
            b = self.terrains[t]
            if b < 0:
                continue
            for a in self.robots[robot]:
                if a >= 0 and b in self.mask[a] and bb_intersect(self.geomList[a][1].getBB(), self.geomList[b][1].getBB()) and self._colliding(a, b):
                    yield self.geomList[a][0], self.geomList[b][0]

    def objectTerrainCollisions(self, object, terrain=None):
        """colliding (RigidObjectModel, TerrainModel) pairs of one object (reference collide.py:632-664: no box pre-reject)"""
        o = object.index if isinstance(object, RigidObjectModel) else object
        ters = range(len(self.terrains)) if terrain is None else [terrain.index if isinstance(terrain, TerrainModel) else terrain]
        a = self.rigidObjects[o]
        for t in ters:
            b = self.terrains[t]
            if a >= 0 and b >= 0 and b in self.mask[a] and self._colliding(a, b):
                yield self.geomList[a][0], self.geomList[b][0]

    def objectObjectCollisions(self, object, object2=None):
        """colliding (object, object2) pairs (reference collide.py:666-698).  With object2 = None the reference calls itself with the
        same arguments and never returns (:687-689); here that case walks over all objects, which is what its comment intends."""
        o = object.index if isinstance(object, RigidObjectModel) else object
        others = range(len(self.rigidObjects)) if object2 is None else [object2.index if isinstance(object2, RigidObjectModel) else object2]
        a = self.rigidObjects[o]
        for o2 in others:
            b = self.rigidObjects[o2]
            if a >= 0 and b >= 0 and a in self.mask[b] and self._colliding(a, b):
                yield self.geomList[a][0], self.geomList[b][0]

    # ------------------------------------------------------------------ ray casts (reference collide.py:700-748)
    def _ray_engine(self):
        """one engine over the whole world (worlds with exactly one robot), rebuilt when the world, a geometry or an object's transform
        changes; None otherwise (callers fall back to one launch per geometry)"""
        w = self.world
        if w.numRobots() != 1:
            return None
        key = (w._version, tuple((g._version, g._T12().tobytes()) for (o, g) in self.geomList if not isinstance(o, RobotModelLink)),
               tuple(g._version for (o, g) in self.geomList if isinstance(o, RobotModelLink)))
        if getattr(self, "_ray_key", None) != key:
            from .engine import Engine
            self._ray_eng, self._ray_key = Engine(w.to_spec()), key
        return self._ray_eng

    def rayCastBatch(self, rays, indices: Optional[List[int]] = None):
        """N rays (rows of source xyz, direction xyz) against the geometries of geomList (all, or those at `indices`) in one launch:
        (geomList index of the nearest hit or -1, distance along the normalised direction or inf, hit points (N, 3))"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        N = len(rays)
        unit = rays[:, 3:] / np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
        eng = self._ray_engine()
        if eng is not None:
            nids = eng.num_ids()
            ignore = None
            if indices is not None:
                ignore = np.ones(nids, dtype=np.uint8)
                ignore[[self._ids[i] for i in indices if i >= 0]] = 0
            ids, dist, _ = eng.raycast_batch(self.world.robot(0).getConfig(), rays, ignore)
            back = np.full(nids + 1, -1, dtype=np.int64)
            back[np.asarray(self._ids, dtype=np.int64)] = np.arange(len(self._ids))
            which = back[ids]                      # id -1 reads the spare last slot
        else:
            which, dist = np.full(N, -1, dtype=np.int64), np.full(N, np.inf)
            for i in (range(len(self.geomList)) if indices is None else indices):
                if i < 0:
                    continue
                el, di = self.geomList[i][1].rayCastBatch(rays)
                better = (el >= 0) & (di < dist)
                which[better], dist[better] = i, di[better]
        pts = rays[:, :3] + np.where(np.isfinite(dist), dist, 0.0)[:, None] * unit
        return which, dist, pts

    def rayCast(self, s, d, indices: Optional[List[int]] = None):
        """first hit of the ray with the world's geometries (or those at `indices`): (object, point) or None"""
        which, dist, pts = self.rayCastBatch(np.concatenate([np.asarray(s, dtype=np.float64), np.asarray(d, dtype=np.float64)])[None, :], indices)
        if which[0] < 0:
            return None
        return self.geomList[int(which[0])][0], list(pts[0])

    def rayCastRobot(self, robot, s, d):
        """first hit of the ray with the links of one robot: (link, point) or None"""
        if isinstance(robot, RobotModel):
            found = [r for r in range(self.world.numRobots()) if self.world.robot(r) is robot]
            if not found:
                raise RuntimeError("Robot " + robot.getName() + " is not found in the world!")
            robot = found[0]
        return self.rayCast(s, d, [i for i in self.robots[robot] if i >= 0])
