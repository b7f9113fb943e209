"""Golden collision masks from the REFERENCE's own Python/klampt/model/collide.py.

``WorldCollider.__init__`` (collide.py:269-362) is pure Python: it only asks the world for terrains / rigid objects / robot links,
their geometry type, ``RobotModelLink.getParent()`` and ``RobotModel.selfCollisionEnabled(i, j)``.  Here it runs UNMODIFIED on this
repo's robotsim mirror objects (registered as ``klampt.robotsim``), so the pair semantics of SURVEY.md 8a rows a6 / a22 -- terrain vs
link only if the link has a parent, object vs everything, self pairs from the robot's matrix, empty geometries skipped,
``ignoreCollision`` edits -- come from the reference's code, not from a restatement.  The masks are committed as
tests/golden/ref_masks.npz; tests/test_reference_golden.py holds this repo's WorldCollider mirror and the oracle's
``InitializeDefault`` mask to them.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_mask.py
"""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


def import_reference_collide():
    from klampt_b200 import robotsim as mirror
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sys.modules["klampt.robotsim"] = mirror                     # `from ..robotsim import *` in collide.py
    return importlib.import_module("klampt.model.collide")


def worlds():
    """name -> WorldSpec; shared with the test (imported from there)"""
    from klampt_b200 import synth
    from klampt_b200.worldspec import GeomSpec
    out = {}
    out["c1"] = synth.world_c1()
    out["c3"] = synth.world_c3()
    w = synth.world_boxes(n_boxes=4, n_blobs=2)
    out["boxes"] = w
    w = synth.world_c1()                                          # an empty link geometry, an empty object, disabled self pairs
    w.robot.link_geom[2] = w.add_geom(GeomSpec("empty"))
    w.objects.append((w.add_geom(GeomSpec("empty")), synth.make_T(None, (0, 0, 0))))
    w.robot.self_collision_edits = [(1, 4, False), (3, 6, False)]
    out["c1_empty"] = w
    return out


def body_key(obj, mirror):
    if isinstance(obj, mirror.TerrainModel):
        return ("terrain", obj.index)
    if isinstance(obj, mirror.RigidObjectModel):
        return ("object", obj.index)
    return ("link", obj.getIndex())


def mask_rows(col, mirror):
    kinds = {"terrain": 0, "object": 1, "link": 2}
    rows = set()
    for i, s in enumerate(col.mask):
        for j in s:
            a, b = body_key(col.geomList[i][0], mirror), body_key(col.geomList[j][0], mirror)
            rows.add((kinds[a[0]], a[1], kinds[b[0]], b[1]))
    return np.array(sorted(rows), dtype=np.int32).reshape(-1, 4)


def main():
    from klampt_b200 import robotsim as mirror
    ref = import_reference_collide()
    out = {}
    for name, spec in worlds().items():
        world = mirror.WorldModel.from_spec(spec)
        col = ref.WorldCollider(world)
        out[name] = mask_rows(col, mirror)
        # ignoreCollision: one whole body, then one pair
        robot = world.robot(0)
        col.ignoreCollision(robot.link(robot.numLinks() - 1))
        if world.numRigidObjects() > 0 and world.rigidObject(0).geometry().type() != "":
            col.ignoreCollision((robot.link(1), world.rigidObject(0)))
        out[name + "_ignored"] = mask_rows(col, mirror)
        print(name, len(out[name]), "->", len(out[name + "_ignored"]), "directed pairs")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_masks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
