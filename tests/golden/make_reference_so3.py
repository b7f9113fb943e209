"""Golden vectors from the REFERENCE's own code: the pure-Python rotation module Python/klampt/math/so3.py (and se3.py) can be
imported without the compiled extension, so its outputs pin the SO(3) part of the oracle -- the geodesic interpolation and angle
metric that Cpp/Modeling/Interpolate.cpp:16-52,229-278 applies to the Euler-ZYX triplets of Floating / BallAndSocket joints, the
roll-pitch-yaw convention of URDF origins and the column-major 9-list convention -- against the real thing instead of a restatement.

Run in the build container (needs /root/reference; the GPU box only reads the .npz):
    python tests/golden/make_reference_so3.py
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


def import_reference_math():
    """klampt/__init__.py imports the compiled robotsim module, which does not exist here: register bare package objects for
    `klampt`, `klampt.math` and `klampt.model` so that only the pure-Python files so3.py / se3.py / vectorops.py / model/typing.py run"""
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    import importlib
    return importlib.import_module("klampt.math.so3"), importlib.import_module("klampt.math.se3"), importlib.import_module("klampt.math.so2")


def main():
    so3, se3, so2 = import_reference_math()
    rng = np.random.default_rng(20261017)
    n = 96
    w = rng.normal(size=(n, 3))
    w *= (rng.uniform(0.0, np.pi, size=n) / np.linalg.norm(w, axis=1))[:, None]
    w[0] = 0.0                                                    # identity
    w[1] = [np.pi - 1e-7, 0, 0]                                   # almost a half turn
    w[2] = [0, 1e-9, 0]                                           # almost nothing
    R = np.array([so3.from_rotation_vector(list(x)) for x in w])                 # 9-lists, column major
    rpy = np.array([so3.rpy(list(r)) for r in R])
    R_from_rpy = np.array([so3.from_rpy(list(x)) for x in rpy])
    moment = np.array([so3.rotation_vector(list(r)) for r in R])
    quat = np.array([so3.quaternion(list(r)) for r in R])
    ang = np.array([so3.angle(list(r)) for r in R])
    ia, ib = rng.integers(0, n, size=200), rng.integers(0, n, size=200)
    u = rng.uniform(0, 1, size=200)
    u[:4] = [0.0, 1.0, 0.5, 0.25]
    interp = np.array([so3.interpolate(list(R[a]), list(R[b]), float(t)) for a, b, t in zip(ia, ib, u)])
    dist = np.array([so3.distance(list(R[a]), list(R[b])) for a, b in zip(ia, ib)])
    mul = np.array([so3.mul(list(R[a]), list(R[b])) for a, b in zip(ia, ib)])
    p = rng.normal(size=(200, 3))
    t = rng.normal(size=(n, 3))
    applied = np.array([se3.apply((list(R[a]), list(t[a])), list(x)) for a, x in zip(ia, p)])
    se3_mul = [se3.mul((list(R[a]), list(t[a])), (list(R[b]), list(t[b]))) for a, b in zip(ia, ib)]
    # so2: the shortest-arc difference / interpolation that Spin joints and the angle of FloatingPlanar joints use
    ang_a, ang_b, ang_u = rng.uniform(-7.0, 7.0, size=300), rng.uniform(-7.0, 7.0, size=300), rng.uniform(0, 1, size=300)
    ang_a[:3], ang_b[:3] = [0.1, 6.2, -3.0], [6.2, 0.1, 3.0]
    so2_diff = np.array([so2.diff(float(a), float(b)) for a, b in zip(ang_a, ang_b)])
    so2_interp = np.array([so2.interp(float(a), float(b), float(t)) for a, b, t in zip(ang_a, ang_b, ang_u)])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_so3.npz")
    np.savez_compressed(out, w=w, R=R, rpy=rpy, R_from_rpy=R_from_rpy, moment=moment, quat=quat, angle=ang, ia=ia, ib=ib, u=u,
                        ang_a=ang_a, ang_b=ang_b, ang_u=ang_u, so2_diff=so2_diff, so2_interp=so2_interp,
                        interp=interp, dist=dist, mul=mul, p=p, t=t, applied=applied,
                        se3_mul_R=np.array([m[0] for m in se3_mul]), se3_mul_t=np.array([m[1] for m in se3_mul]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
