"""Batched form of the temporal-coherence distance query.

Reference: ``DistanceQuery`` (Cpp/Planning/DistanceQuery.{h,cpp}): a per-pair state machine Far / Close / Contact that
chooses the cheapest sufficient query for the next cycle -- ``WithinDistance(distanceTolerance)`` while the pair was far,
``Distance(absErr, relErr)`` while it was close, ``PenetrationDepth()`` while it was in contact -- and reports
``distanceTolerance`` for far pairs, the distance for close pairs and minus the penetration depth for contacts
(DistanceQuery.cpp:25-76).

Here one object tracks N transform pairs of the same two geometries (N trajectories, N particles ...) and every cycle
runs at most two launches over the subsets that need them: the boolean within-distance kernel on the pairs that were far,
the branch-and-bound distance kernel (exact: absErr = relErr = 0) on the rest.  Penetration depth is not computed by the
engine (neither is it for mesh pairs in the reference's PQP back end): a pair in contact reports ``-0.0`` and
``penetration_supported`` is False.
"""
from __future__ import annotations

import numpy as np

from .robotsim import Geometry3D

FAR, CLOSE, CONTACT, WAS_FAR, WAS_CLOSE, WAS_CONTACT, INVALID = range(7)
_DEFAULT_TOLERANCE, _DEFAULT_ABS_ERR, _DEFAULT_REL_ERR = 0.2, 0.05, 0.1      # DistanceQuery.cpp:5


class DistanceQueryBatch:
    penetration_supported = False

    def __init__(self, a: Geometry3D, b: Geometry3D, n: int):
        self.a, self.b, self.n = a, b, int(n)
        self.s = np.full(self.n, INVALID, dtype=np.uint8)
        self.distanceTolerance = _DEFAULT_TOLERANCE
        self.distanceAbsErr = _DEFAULT_ABS_ERR       # kept for interface parity; the kernels are exact
        self.distanceRelErr = _DEFAULT_REL_ERR
        self._d = np.full(self.n, np.nan)
        self._eng, self._ga, self._gb = a._pair_engine(b)
        self.launches = 0

    def NextCycle(self):
        """DistanceQuery::NextCycle (DistanceQuery.cpp:15-23) for every pair"""
        m = self.s <= CONTACT
        self.s[m] += 3                               # Far -> WasFar, Close -> WasClose, Contact -> WasContact

    def UpdateQuery(self, Ta, Tb) -> np.ndarray:
        """DistanceQuery::UpdateQuery (DistanceQuery.cpp:25-76) at N transform pairs (row-major 12-vectors): returns the
        separation per pair -- distanceTolerance if far, the distance if close, -0.0 (no penetration depth) in contact."""
        Ta = np.ascontiguousarray(Ta, dtype=np.float64).reshape(self.n, 12)
        Tb = np.ascontiguousarray(Tb, dtype=np.float64).reshape(self.n, 12)
        tol = self.distanceTolerance
        out = np.empty(self.n)
        done = self.s <= CONTACT                     # already evaluated this cycle: repeat the cached answer
        out[done] = self._d[done]
        was_far = (self.s == WAS_FAR) | (self.s == INVALID)
        need_d = (self.s == WAS_CLOSE) | (self.s == WAS_CONTACT)
        idx = np.nonzero(was_far)[0]
        if len(idx):                                  # far pairs: the cheap boolean query first
            within = self._eng.geom_collides_batch(self._ga, Ta[idx], self._gb, Tb[idx], tol=tol).astype(bool)
            self.launches += 1
            far = idx[~within]
            self.s[far] = FAR
            out[far] = tol
            need_d[idx[within]] = True
        idx = np.nonzero(need_d)[0]
        if len(idx):
            d = self._eng.geom_distance_batch(self._ga, Ta[idx], self._gb, Tb[idx], upper_bound=tol)
            self.launches += 1
            contact = d <= 0.0
            close = (~contact) & (d < tol)
            far = (~contact) & (~close)
            self.s[idx[contact]] = CONTACT
            self.s[idx[close]] = CLOSE
            self.s[idx[far]] = FAR
            out[idx[contact]] = -0.0
            out[idx[close]] = d[close]
            out[idx[far]] = tol
        self._d = out.copy()
        return out
