"""Golden outputs of the REFERENCE's WorldCollider iterators (Python/klampt/model/collide.py:429-698: collisionTests, collisions,
robotSelfCollisions, robotObjectCollisions, robotTerrainCollisions, objectTerrainCollisions, objectObjectCollisions), run unmodified
on this repo's mirror worlds with every Geometry3D's getBB / collides replaced by fixed pseudo-random boxes (box overlap) -- what is
pinned is WHICH pairs each iterator visits and reports, and in which order and orientation, not the geometry kernel.
-> tests/golden/ref_iterators.json

Run in the build container (needs /root/reference):   python tests/golden/make_reference_iterators.py
"""
import json
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def body_key(obj):
    return [type(obj).__name__, int(obj.index)]


def stub_geometries(world, seed):
    """fixed box per body, keyed by (kind, index): getBB is the box grown by 2 cm, collides is overlap of the boxes themselves"""
    def box(kind, idx):
        rng = np.random.default_rng(zlib.crc32(("%s/%d/%d" % (kind, idx, seed)).encode()))
        c, h = rng.uniform(-0.6, 0.6, 3), rng.uniform(0.1, 0.45, 3)
        return list(c - h), list(c + h)
    bodies = [world.terrain(i) for i in range(world.numTerrains())] + [world.rigidObject(i) for i in range(world.numRigidObjects())]
    bodies += [world.robot(0).link(j) for j in range(world.robot(0).numLinks())]
    for b in bodies:
        g = b.geometry()
        g._stub = box(type(b).__name__, b.index)
        g.getBB = (lambda g=g: ([x - 0.02 for x in g._stub[0]], [x + 0.02 for x in g._stub[1]]))
        g.collides = (lambda o, g=g: not any(q < u or v < p for p, q, u, v in zip(g._stub[0], g._stub[1], o._stub[0], o._stub[1])))


def is_link(body):
    return type(body).__name__ == "RobotModelLink"


def is_env(body):
    return type(body).__name__ != "RobotModelLink"


def run_iterators(col, world):
    P = lambda it: [[body_key(a), body_key(b)] for a, b in it]
    out = {"collisionTests": [[body_key(A[0]), body_key(B[0])] for A, B in col.collisionTests()],
           "collisionTests_nobb": len(list(col.collisionTests(bb_reject=False))),
           "collisions": P(col.collisions()),
           "collisionTests_links": [[body_key(A[0]), body_key(B[0])] for A, B in col.collisionTests(is_link)],
           "collisionTests_links_vs_env": [[body_key(A[0]), body_key(B[0])] for A, B in col.collisionTests(is_link, is_env)],
           "collisions_env_vs_links": P(col.collisions(is_env, is_link)),
           "robotSelfCollisions": P(col.robotSelfCollisions(0)),
           "robotObjectCollisions": P(col.robotObjectCollisions(0)),
           "robotTerrainCollisions": P(col.robotTerrainCollisions(0))}
    if world.numRigidObjects() > 1:
        out["robotObjectCollisions_1"] = P(col.robotObjectCollisions(0, 1))
        # with object2 = None the reference recurses into itself forever (collide.py:687-689): explicit pairs only
        n = min(world.numRigidObjects(), 6)
        out["objectObjectCollisions"] = [[i, j, P(col.objectObjectCollisions(i, j))] for i in range(n) for j in range(n)]
    if world.numRigidObjects() > 0 and world.numTerrains() > 0:
        out["objectTerrainCollisions"] = [[i, P(col.objectTerrainCollisions(i))] for i in range(min(world.numRigidObjects(), 6))]
        out["objectTerrainCollisions_0_0"] = P(col.objectTerrainCollisions(0, 0))
    return out


def main():
    from make_reference_mask import import_reference_collide, worlds
    from klampt_b200 import robotsim as mirror
    ref = import_reference_collide()
    out = {}
    for name, spec in worlds().items():
        for seed in (0, 1):
            world = mirror.WorldModel.from_spec(spec)
            stub_geometries(world, seed)
            out["%s/%d" % (name, seed)] = run_iterators(ref.WorldCollider(world), world)
            print(name, seed, {k: (len(v) if isinstance(v, list) else v) for k, v in out["%s/%d" % (name, seed)].items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_iterators.json")
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
