"""A .rob file written by the REFERENCE's own generator, Python/klampt/model/create/planar_robot.py:3-77 (pure Python up to the
``world.loadElement`` call, which a stub world intercepts): tests/golden/ref_planar_3R.rob is byte for byte what Klamp't writes and
then loads.  tests/test_reference_golden.py feeds it to klampt_b200.io.parse_rob.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_rob.py
"""
import importlib
import importlib.util
import os
import sys
import tempfile
import types

REF = os.environ.get("KLAMPT_REFERENCE", "/root/reference")


class _StubWorld:
    """captures the file the generator hands to WorldModel.loadElement; everything after that needs the compiled module"""
    text = None

    def loadElement(self, fn):
        self.text = open(fn).read()
        raise _Captured()


class _Captured(Exception):
    pass


def main():
    spec = importlib.util.spec_from_file_location("planar_robot", os.path.join(REF, "Python", "klampt", "model", "create", "planar_robot.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    here = os.path.dirname(os.path.abspath(__file__))
    for n, length in ((3, 0.5),):
        w = _StubWorld()
        tmp = os.path.join(tempfile.mkdtemp(), "temp.rob")
        try:
            mod.make(n, w, link_length=length, tempname=tmp, debug=True)
        except _Captured:
            pass
        out = os.path.join(here, "ref_planar_%dR.rob" % n)
        open(out, "w").write(w.text)
        print("wrote", out, len(w.text), "bytes")
    # the free-floating base around a geometry file: model/create/moving_base_robot.py:12-137 (needs `klampt` names at import: stubs)
    root = os.path.join(REF, "Python", "klampt")
    for name, path in (("klampt", root), ("klampt.math", os.path.join(root, "math")), ("klampt.model", os.path.join(root, "model")),
                       ("klampt.model.create", os.path.join(root, "model", "create"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    for cls in ("WorldModel", "RobotModel", "SimRobotController"):
        setattr(sys.modules["klampt"], cls, type(cls, (), {}))
    mb = importlib.import_module("klampt.model.create.moving_base_robot")
    w = _StubWorld()
    tmp = os.path.join(tempfile.mkdtemp(), "temp.rob")
    try:
        mb.make("cube.off", w, tempname=tmp, debug=True)
    except _Captured:
        pass
    out = os.path.join(here, "ref_moving_base.rob")
    open(out, "w").write(w.text)
    print("wrote", out, len(w.text), "bytes")
    # ... and around a robot file: the template with the `mount 5 "<file>" <T> as "<name>"` line
    w = _StubWorld()
    try:
        mb.make("ref_planar_3R.rob", w, tempname=tmp, debug=True)
    except _Captured:
        pass
    out = os.path.join(here, "ref_moving_base_mounted.rob")
    open(out, "w").write(w.text)
    print("wrote", out, len(w.text), "bytes")


if __name__ == "__main__":
    main()
