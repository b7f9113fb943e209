"""Golden feasibility-test lists from the REFERENCE's own Python/klampt/plan/robotcspace.py + plan/cspace.py.

``RobotCSpace.__init__`` (robotcspace.py:31-75) is pure Python.  It runs here UNMODIFIED on this repo's robotsim / collide mirror
objects (the compiled ``motionplanning`` module is only needed by ``setup()``, which is not called): the names, order and
dependencies of the feasibility tests, the bounds and the properties it derives are the reference's, and
tests/test_reference_golden.py holds klampt_b200.robotcspace.RobotCSpace to them.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_cspace.py
"""
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


def import_reference_robotcspace():
    from klampt_b200 import robotsim as mirror
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model")),
                       ("klampt.plan", os.path.join(root, "plan"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    rs = types.ModuleType("klampt.robotsim")                    # the mirror's names + the one extra name robotcspace.py imports
    rs.__dict__.update({k: v for k, v in mirror.__dict__.items() if not k.startswith("__")})
    rs.IKObjective = type("IKObjective", (), {})
    sys.modules["klampt.robotsim"] = rs
    sys.modules["klampt"].robotsim = rs
    sys.modules["klampt.plan.motionplanning"] = types.ModuleType("klampt.plan.motionplanning")     # SWIG module: only setup() uses it
    sys.modules["klampt.plan"].motionplanning = sys.modules["klampt.plan.motionplanning"]
    collide = importlib.import_module("klampt.model.collide")
    return importlib.import_module("klampt.plan.robotcspace"), collide


def main():
    from make_reference_mask import worlds
    from klampt_b200 import robotsim as mirror
    ref, collide = import_reference_robotcspace()
    out = {}
    for name, spec in worlds().items():
        for with_collider in (True, False):
            world = mirror.WorldModel.from_spec(spec)
            robot = world.robot(0)
            space = ref.RobotCSpace(robot, collide.WorldCollider(world) if with_collider else None)
            out[name + ("" if with_collider else "_nocollider")] = {
                "names": list(space.feasibilityTestNames),
                "dependencies": [list(d) for d in space.feasibilityTestDependencies],
                "bound": [[float(a), float(b)] for a, b in space.bound],
                "properties": {k: (v if not isinstance(v, (list, tuple)) else [float(x) for x in v]) for k, v in space.properties.items()},
                "eps": space.eps,
                "in_limits": [bool(space.inJointLimits([b[0] for b in space.bound])), bool(space.inJointLimits([b[1] + 1e-9 for b in space.bound]))],
            }
            print(name, with_collider, len(space.feasibilityTestNames), "tests")
    # EmbeddedRobotCSpace.disableInactiveCollisions (plan/robotcspace.py:365-392) on the arm: which mask rows it rewrites
    from make_reference_mask import mask_rows
    for subset in ([1, 2], [4, 5, 6], [6], [0, 3]):
        world = mirror.WorldModel.from_spec(worlds()["c1"])
        robot = world.robot(0)
        robot._q = __import__("numpy").zeros(robot.numLinks())
        col = collide.WorldCollider(world)
        sp = ref.EmbeddedRobotCSpace(ref.RobotCSpace(robot, col), subset, xinit=None)
        sp.disableInactiveCollisions()
        out["inactive_" + "_".join(map(str, subset))] = {"mask": mask_rows(col, mirror).tolist(), "xinit": list(map(float, sp.xinit)), "bound": [list(b) for b in sp.bound]}
    # the plain CSpace base class (plan/cspace.py:76-214), members that do not need the compiled CSpaceInterface
    base = importlib.import_module("klampt.plan.cspace").CSpace()
    base.setBounds([(0.0, 2.0), (1.0, 1.0), (-1.0, 3.0)])
    base.addFeasibilityTest(lambda x: x[0] < 1.5)
    base.addFeasibilityTest(lambda x: x[2] > 0.0, "positive z", dependencies=["test_0"])
    base.addFeasibilityTest(lambda x: True, dependencies="positive z")
    probes = [[0.5, 1.0, 1.0], [1.7, 1.0, 1.0], [0.5, 1.0, -0.5], [0.5, 1.2, 1.0], [2.0, 1.0, 3.0]]
    out["plain_cspace"] = {"bound": [list(b) for b in base.bound], "properties": dict(base.properties), "eps": base.eps,
                           "names": list(base.feasibilityTestNames), "dependencies": [list(d) for d in base.feasibilityTestDependencies],
                           "probes": probes, "inBounds": [bool(base.inBounds(p)) for p in probes], "feasible": [bool(base.feasible(p)) for p in probes],
                           "stats_before_setup": base.getStats()}
    # EmbeddedCSpace (plan/cspaceutils.py:108-203) over that plain space: DOFs 2 and 0 move, DOF 1 stays at xinit
    utils = importlib.import_module("klampt.plan.cspaceutils")
    base.distance = lambda a, b: sum(abs(p - q) for p, q in zip(a, b))                    # optional methods: a weighted-free L1 metric here
    base.interpolate = lambda a, b, u: [p + u * (q - p) for p, q in zip(a, b)]
    emb = utils.EmbeddedCSpace(base, [2, 0], xinit=[0.25, 1.0, 0.5])
    eprobes = [[1.0, 0.5], [-0.5, 0.5], [1.0, 1.7], [3.0, 2.0]]
    out["embedded_cspace"] = {"bound": [list(b) for b in emb.bound], "eps": emb.eps, "names": list(emb.feasibilityTestNames),
                              "dependencies": [list(d) for d in emb.feasibilityTestDependencies], "probes": eprobes,
                              "lift": [emb.lift(p) for p in eprobes], "project": emb.project([9.0, 8.0, 7.0]),
                              "feasible": [bool(emb.feasible(p)) for p in eprobes],
                              "tests": [[bool(f(p)) for f in emb.feasibilityTests] for p in eprobes],
                              "distance": emb.distance(eprobes[0], eprobes[2]), "interpolate": emb.interpolate(eprobes[0], eprobes[2], 0.25),
                              "liftPath": emb.liftPath(eprobes[:2]), "projectPath": emb.projectPath([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]),
                              "default_xinit_lift": utils.EmbeddedCSpace(base, [1]).lift([0.75])}
    # AffineEmbeddedCSpace (plan/cspaceutils.py:205-421): driver space of a 4-link chain -- link 0 alone, links 1 and 2 coupled
    # (q1 = 2 v, q2 = -v + 0.1), link 3 not driven (stays at b)
    import numpy as _np, io as _io, contextlib as _ctx
    amb = importlib.import_module("klampt.plan.cspace").CSpace()
    amb.setBounds([(-1.0, 1.0), (-2.0, 2.0), (-0.5, 0.5), (0.0, 1.0)])
    A = _np.array([[1.0, 0.0], [0.0, 2.0], [0.0, -1.0], [0.0, 0.0]])
    boff = [0.0, 0.0, 0.1, 0.25]
    with _ctx.redirect_stdout(_io.StringIO()):
        aff = utils.AffineEmbeddedCSpace(amb, A, boff)
    xs = [[0.5, 0.2], [-1.0, -0.4], [0.0, 0.0]]
    out["affine_cspace"] = {"lift": [list(map(float, aff.lift(x))) for x in xs],
                            "project": [list(map(float, aff.project(aff.lift(x)))) for x in xs],
                            "project_off_manifold": list(map(float, aff.project([0.3, 1.0, 0.2, 0.9]))),
                            "bound_as_sets": [sorted(map(float, bd)) for bd in aff.bound], "eps": aff.eps}
    # AffineEmbeddedCSpace.fromRobotDrivers (:320-369) on a mirror robot whose last two links share an affine driver
    from klampt_b200.worldspec import DriverSpec
    spec = worlds()["c1"]
    spec.robot.drivers = [DriverSpec([k], [1.0], [0.0], -2.0, 2.0) for k in range(1, 5)] + [DriverSpec([5, 6], [1.0, -0.5], [0.0, 0.2], -1.0, 1.0)]
    world = mirror.WorldModel.from_spec(spec)
    robot = world.robot(0)
    robot._q = _np.zeros(robot.numLinks())
    with _ctx.redirect_stdout(_io.StringIO()):
        fr = utils.AffineEmbeddedCSpace.fromRobotDrivers(robot, ref.RobotCSpace(robot, None))
    out["affine_from_drivers"] = {"A": _np.asarray(fr.A.todense() if hasattr(fr.A, "todense") else fr.A).tolist(), "b": list(map(float, fr.b)),
                                  "lift": list(map(float, fr.lift([0.1, 0.2, 0.3, 0.4, 0.5])))}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cspace.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
