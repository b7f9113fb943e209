"""Golden forward kinematics from the REFERENCE's own code: ``KinematicsBuilder.__init__`` in
Python/klampt/math/autodiff/kinematics_ad.py:407-457 is the in-repo restatement of RobotKinematics3D::UpdateFrames (SURVEY.md 8a
row a3).  With a fixed configuration it evaluates numerically, and it only asks the robot for getConfig / numLinks /
link(i).getParent / getParentTransform / getAxis / isPrismatic -- all of which this repo's robotsim mirror answers.  So the
reference's FK code runs here UNMODIFIED on the procedural robots, and tests/golden/ref_fk.npz (configurations + every link's
world transform) pins the oracle's FK recurrence (and, through the GPU parity tests, the FK kernel) to the reference.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_fk.py
"""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


def import_reference_kinematics():
    from klampt_b200 import robotsim as mirror
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model")),
                       ("klampt.math.autodiff", os.path.join(root, "math", "autodiff"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sys.modules["klampt.robotsim"] = mirror
    return importlib.import_module("klampt.math.autodiff.kinematics_ad")


def robots():
    """name -> (WorldSpec, seed); the test rebuilds the same specs"""
    from klampt_b200 import synth
    from klampt_b200.worldspec import WorldSpec
    out = {"arm6": synth.world_c1(), "dualarm15": synth.world_c3(), "floating": synth.world_floating()}
    w = WorldSpec(); w.robot = synth.make_planar_nR(w, 5, 0.4); out["planar5R"] = w
    return out


def to12(T):
    """reference transform (column-major 9-list R, t) or flat 12-array (R column-major, then t) -> the C ABI's row-major 12-vector"""
    if hasattr(T, "eval"):                # an expression of constants (the configuration is fixed): evaluate it
        T = T.eval()
    a = np.asarray(T[0] + T[1] if isinstance(T, tuple) else T, dtype=np.float64).reshape(-1)
    return np.concatenate([a[:9].reshape(3, 3).T.reshape(-1), a[9:12]])


def main():
    from klampt_b200 import robotsim as mirror, synth
    kin = import_reference_kinematics()
    out = {}
    for name, spec in robots().items():
        world = mirror.WorldModel.from_spec(spec)
        robot = world.robot(0)
        Q = synth.sample_configs(spec.robot, 24, 77)
        Q[0] = 0.0
        T = np.zeros((len(Q), spec.robot.L, 12))
        for k, q in enumerate(Q):
            robot._q = np.asarray(q, dtype=np.float64)            # the configuration only: the reference's code does the kinematics
            kb = kin.KinematicsBuilder(robot)
            for i in range(spec.robot.L):
                T[k, i] = to12(kb.link_transforms[i])
        out[name + "_Q"], out[name + "_T"] = Q, T
        print(name, Q.shape, T.shape)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_fk.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
