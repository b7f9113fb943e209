"""Golden outputs of the REFERENCE's module-level helpers in Python/klampt/model/collide.py (bb_* and the group collision iterators,
:24-216), run unmodified on stand-in geometries (axis-aligned boxes whose `collides` is box overlap) -> tests/golden/ref_groupiter.json.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_groupiter.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class BoxGeom:
    """what the iterators need of a Geometry3D: getBB() and collides()"""
    def __init__(self, lo, hi, loose=0.0):
        self.lo, self.hi, self.loose = [float(x) for x in lo], [float(x) for x in hi], loose

    def getBB(self):
        return [a - self.loose for a in self.lo], [b + self.loose for b in self.hi]

    def collides(self, o):
        return not any(q < u or v < p for p, q, u, v in zip(self.lo, self.hi, o.lo, o.hi))


def boxes(seed, n, spread):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-spread, spread, size=(n, 3))
    h = rng.uniform(0.05, 0.4, size=(n, 3))
    return [(list(a - b), list(a + b)) for a, b in zip(c, h)]


CASES = {"dense": (1, 28, 1.0), "sparse": (2, 14, 4.0), "far_groups": (3, 16, 0.8)}


def case_inputs(name):
    seed, n, spread = CASES[name]
    b = boxes(seed, n, spread)
    g1, g2 = b[:n // 2], b[n // 2:]
    if name == "far_groups":
        g2 = [([x + 10 for x in lo], [x + 10 for x in hi]) for lo, hi in g2[:-1]] + [g2[-1]]
    return b, g1, g2


def main():
    from make_reference_mask import import_reference_collide
    ref = import_reference_collide()
    out = {}
    for name in CASES:
        b, g1, g2 = case_inputs(name)
        G, G1, G2 = [BoxGeom(*x, loose=0.01) for x in b], [BoxGeom(*x, loose=0.01) for x in g1], [BoxGeom(*x, loose=0.01) for x in g2]
        even = lambda i, j: (i + j) % 2 == 0
        sub_a, sub_b = list(range(0, len(G), 2)) + [1], [0, 2, 4, 1]          # blist inside alist: the only case the reference handles
        out[name] = {
            "self_all": [list(p) for p in ref.self_collision_iter(G)],
            "self_even": [list(p) for p in ref.self_collision_iter(G, even)],
            "self_list": [list(p) for p in ref.self_collision_iter(G, [(0, 1), (2, 5), (3, 4)])],
            "group_all": [list(p) for p in ref.group_collision_iter(G1, G2)],
            "group_even": [list(p) for p in ref.group_collision_iter(G1, G2, even)],
            "group_list": [list(p) for p in ref.group_collision_iter(G1, G2, [(0, 0), (1, 2)])],
            "subset_all": [list(p) for p in ref.group_subset_collision_iter(G, sub_a, sub_b)],
            "subset": [sub_a, sub_b],
            "bb_union": list(map(list, ref.bb_union(*[g.getBB() for g in G]))),
            "bb_intersection": list(map(list, ref.bb_intersection(G[0].getBB(), G[1].getBB()))),
            "bb_empty": bool(ref.bb_empty(ref.bb_intersection(G[0].getBB(), G[1].getBB()))),
            "bb_create": list(map(list, ref.bb_create(*[g.lo for g in G]))),
            "bb_contains": [bool(ref.bb_contains(G[0].getBB(), g.lo)) for g in G],
        }
    out["bb_create_empty"] = list(map(list, ref.bb_create()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_groupiter.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, {k: len(v["self_all"]) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
