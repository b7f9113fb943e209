"""Ingestion formats (OFF / OBJ meshes, .rob robots) -- host only."""
import math
import os

import numpy as np
import pytest

from klampt_b200 import io as kio
from klampt_b200 import synth
from klampt_b200.worldspec import WorldSpec, JOINT_SPIN, JOINT_WELD, JOINT_NORMAL
from oracle.oracle import OracleWorld

CUBE_OFF = """OFF
8 12 0
0.000000 0.000000 0.000000
0.000000 0.000000 1.000000
0.000000 1.000000 0.000000
0.000000 1.000000 1.000000
1.000000 0.000000 0.000000
1.000000 0.000000 1.000000
1.000000 1.000000 0.000000
1.000000 1.000000 1.000000
3 0 1 3
3 0 3 2
3 4 6 7
3 4 7 5
3 0 4 5
3 0 5 1
3 2 3 7
3 2 7 6
3 0 2 6
3 0 6 4
3 1 5 7
3 1 7 3
"""


def test_off_matches_the_procedural_unit_cube(tmp_path):
    v, t = kio.parse_off(CUBE_OFF)
    v0, t0 = synth.unit_cube()
    assert np.array_equal(v, v0) and np.array_equal(t, t0)
    p = tmp_path / "c.off"
    kio.save_off(str(p), v, t)
    v2, t2 = kio.load_mesh(str(p))
    assert np.array_equal(v2, v) and np.array_equal(t2, t)
    q = tmp_path / "quad.obj"
    q.write_text("# quad\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\nf -4/1 -3/2 -2/3\n")
    vq, tq = kio.load_mesh(str(q))
    assert len(vq) == 4 and tq.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    with pytest.raises(ValueError):
        kio.parse_off("OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 7\n")


def reference_style_planar_rob(n, link_length=0.5):
    """the text Python/klampt/model/create/planar_robot.py writes (same keywords, continuation lines, inline OFF geometry)"""
    geom = ('"{TriangleMesh\\nOFF\\n8 12 0\\n0.0 -0.05 -0.05\\n0.0 -0.05 0.05\\n0.0 0.05 -0.05\\n0.0 0.05 0.05\\n1.0 -0.05 -0.05\\n1.0 -0.05 0.05\\n'
            '1.0 0.05 -0.05\\n1.0 0.05 0.05\\n3 0 1 3\\n3 0 3 2\\n3 4 6 7\\n3 4 7 5\\n3 0 4 5\\n3 0 5 1\\n3 2 3 7\\n3 2 7 6\\n3 0 2 6\\n3 0 6 4\\n3 1 5 7\\n3 1 7 3\\n}"')
    s = "### A %dR planar robot ###\nTParent 1 0 0   0 1 0   0 0 1   0 0 0 " % n
    s += ("   \\\n1 0 0   0 1 0   0 0 1   %g 0 0" % link_length) * (n - 1) + "\n"
    s += "axis\t" + "\t".join(["0 1 0"] * n) + "\n" + "jointtype\t" + " ".join(["r"] * n) + "\n"
    s += "qMin\t" + " ".join(["0"] * n) + "\nqMax\t" + " ".join(["6.28319"] * n) + "\nq\t" + " ".join(["0"] * n) + "\n"
    s += "geometry\t" + " ".join([geom] * n) + "\ngeomscale\t" + " ".join(["%g" % max(link_length, 0.05)] * n) + "\n"
    s += "mass\t" + " ".join(["1"] * n) + "\nautomass\ntorqueMax\t" + " ".join(["10"] * n) + "\naccMax\t" + " ".join(["1"] * n) + "\nvelMax\t" + " ".join(["1"] * n) + "\n"
    s += "parents " + " ".join(str(i - 1) for i in range(n)) + "\n"
    for i in range(n):
        s += "joint spin %d\n" % i
    s += "servoP\t" + " ".join(["50"] * n) + "\n"
    return s


def test_rob_loader_on_reference_style_planar_robot():
    world, r = kio.parse_rob(reference_style_planar_rob(4, 0.5))
    assert r.L == 4 and list(r.parents) == [-1, 0, 1, 2]
    assert np.allclose(r.T0[1:, 9], 0.5) and np.allclose(r.T0[0, 9:], 0) and np.allclose(r.axis, [[0, 1, 0]] * 4)
    assert (r.joint_type == JOINT_SPIN).all() and list(r.joint_link) == [0, 1, 2, 3]
    assert np.allclose(r.qmax, 6.28319) and all(g >= 0 for g in r.link_geom)
    g = world.geoms[r.link_geom[2]]
    assert g.tris.shape == (12, 3) and g.verts[:, 0].max() == pytest.approx(0.5)          # geomscale applied
    # the loaded robot drives the oracle: closed-form FK of the planar chain
    o = OracleWorld(world)
    q = [0.3, -0.2, 0.5, 0.1]
    T = o.fk(q)
    th = 0.3 - 0.2
    assert np.allclose(T[2, 9:], [0.5 * math.cos(0.3) + 0.5 * math.cos(th), 0, -0.5 * math.sin(0.3) - 0.5 * math.sin(th)], atol=1e-12)


def test_rob_round_trip_of_the_procedural_arm():
    w = synth.world_c1()
    text = kio.rob_text(w.robot, w)
    w2, r2 = kio.parse_rob(text)
    r = w.robot
    assert r2.names == r.names and np.array_equal(r2.parents, r.parents) and np.array_equal(r2.linktype, r.linktype)
    assert np.array_equal(r2.T0, r.T0) and np.array_equal(r2.axis, r.axis) and np.array_equal(r2.qmin, r.qmin) and np.array_equal(r2.qmax, r.qmax)
    assert list(r2.joint_type) == [JOINT_WELD] + [JOINT_NORMAL] * 6
    for a, b in zip(r.link_geom, r2.link_geom):
        assert np.array_equal(w.geoms[a].verts, w2.geoms[b].verts) and np.array_equal(w.geoms[a].tris, w2.geoms[b].tris)
    # same feasibility answers from the re-loaded robot in the same environment
    w2.terrains, w2.objects = [], []
    for gi in w.terrains:
        w2.terrains.append(w2.add_geom(w.geoms[gi]))
    for gi, T in w.objects:
        w2.objects.append((w2.add_geom(w.geoms[gi]), T))
    Q = synth.sample_configs(r, 300, 3)
    assert np.array_equal(OracleWorld(w).feasible_batch(Q), OracleWorld(w2).feasible_batch(Q))


def test_rob_base_transform_selfcollision_and_errors():
    text = ("links a b c\nparents -1 0 1\ntparent 1 0 0 0 1 0 0 0 1 0 0 0  1 0 0 0 1 0 0 0 1 1 0 0  1 0 0 0 1 0 0 0 1 1 0 0\n"
            "qmindeg -90 -90 -90\nqmaxdeg 90 90 90\ntranslation 0 0 2\nrotation 0 -1 0 1 0 0 0 0 1\nnoselfcollision a c\njoint weld 0\njoint normal 1\njoint normal 2\n"
            "driver normal 1\ndriver affine 2 1 2 1.0 -1.0 0 0 -0.5 0.5\n")
    world, r = kio.parse_rob(text)
    assert np.allclose(r.qmin, -math.pi / 2) and np.allclose(r.T0[0, 9:], [0, 0, 2]) and np.allclose(r.T0[0, :9], [0, -1, 0, 1, 0, 0, 0, 0, 1])
    assert np.allclose(r.T0[1, 9:], [1, 0, 0])                       # only root links take the base transform
    assert (0, 2, False) in r.self_collision_edits and r.joint_type[0] == JOINT_WELD
    assert len(r.drivers) == 2 and r.drivers[1].links == [1, 2] and r.drivers[1].scale == [1.0, -1.0] and r.drivers[1].qmax == 0.5
    with pytest.raises(NotImplementedError):
        kio.parse_rob("parents -1\nalpha 0\n")
    with pytest.raises(ValueError):
        kio.parse_rob("links a\n")
