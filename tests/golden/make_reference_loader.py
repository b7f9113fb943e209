"""Golden text of the REFERENCE's own resource formats for configurations and transforms: Python/klampt/io/loader.py
(write_Vector / write_VectorList / write_se3 and their readers, :161-198,226-250) -- the .config / .configs / .xform files that feed
batches of configurations and object poses to the path.  -> tests/golden/ref_loader.json

Run in the build container (needs /root/reference):   python tests/golden/make_reference_loader.py
"""
import importlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


def import_reference_loader():
    """loader.py imports contact / trajectory / types modules that need the compiled extension; only its text functions are wanted,
    so those three are registered as empty modules carrying the names it imports"""
    from klampt_b200 import robotsim as mirror
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model")),
                       ("klampt.io", os.path.join(root, "io"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sys.modules["klampt.robotsim"] = mirror
    for name, attrs in (("klampt.model.contact", ("ContactPoint", "Hold")),
                        ("klampt.model.trajectory", ("Trajectory", "HermiteTrajectory", "SO3Trajectory", "SE3Trajectory")), ("klampt.model.types", ())):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        sys.modules[name] = m
        setattr(sys.modules["klampt.model"], name.rsplit(".", 1)[1], m)
    return importlib.import_module("klampt.io.loader"), importlib.import_module("klampt.math.so3")


def main():
    loader, so3 = import_reference_loader()
    rng = np.random.default_rng(20261017)
    Q = rng.uniform(-3, 3, size=(5, 7))
    Q[0, 0], Q[1, 1], Q[2, 2] = 0.0, 1e-17, -123456.789
    R = so3.from_rotation_vector([0.3, -0.5, 0.8])
    t = [0.25, -1.5, 3.0]
    out = {"Q": Q.tolist(), "config_text": loader.write_Vector(list(Q[0])), "configs_text": loader.write_VectorList([list(q) for q in Q]),
           "R": list(R), "t": t, "xform_text": loader.write_se3((R, t)),
           "read_back_configs": loader.read_VectorList(loader.write_VectorList([list(q) for q in Q])),
           "read_back_xform": [list(x) for x in loader.read_se3(loader.write_se3((R, t)))]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_loader.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path); print(out["config_text"]); print(out["xform_text"])


if __name__ == "__main__":
    main()
