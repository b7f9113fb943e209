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


URDF_ARM = """<robot name="two">
 <link name="world"/>
 <link name="base"><collision><origin xyz="0 0 0.05"/><geometry><box size="0.2 0.2 0.1"/></geometry></collision></link>
 <link name="arm"><collision><origin xyz="0 0 0.25"/><geometry><cylinder radius="0.04" length="0.5"/></geometry></collision></link>
 <link name="finger"><collision><geometry><sphere radius="0.03"/></geometry></collision></link>
 <link name="finger2"/>
 <joint name="fix" type="fixed"><parent link="world"/><child link="base"/><origin xyz="0 0 0.1"/></joint>
 <joint name="j1" type="revolute"><parent link="base"/><child link="arm"/><origin xyz="0 0 0.1" rpy="0 0 1.5707963267948966"/><axis xyz="0 1 0"/><limit lower="-1" upper="2"/></joint>
 <joint name="j2" type="prismatic"><parent link="arm"/><child link="finger"/><origin xyz="0 0 0.5"/><axis xyz="0 0 1"/><limit lower="0" upper="0.1"/></joint>
 <joint name="j3" type="continuous"><parent link="arm"/><child link="finger2"/><origin xyz="0 0.1 0.5"/><axis xyz="0 0 2"/><mimic joint="j1" multiplier="2" offset="0.1"/></joint>
 <klampt><noselfcollision pairs="base finger"/></klampt>
</robot>"""


def test_urdf_fixed_base_follows_the_reference_loader():
    """RobotModel::LoadURDF (reference Cpp/Modeling/Robot.cpp:2864-3135): the world link is dropped, joint types, limits,
    mimic -> affine driver, <klampt> self-collision edits; FK of the result through the oracle."""
    from klampt_b200.worldspec import PRISMATIC, REVOLUTE
    w, r = kio.parse_urdf(URDF_ARM)
    assert r.names == ["base", "arm", "finger", "finger2"] and list(r.parents) == [-1, 0, 1, 1]
    assert list(r.linktype) == [REVOLUTE, REVOLUTE, PRISMATIC, REVOLUTE]
    assert list(r.joint_type) == [JOINT_WELD, JOINT_NORMAL, JOINT_NORMAL, JOINT_SPIN]
    assert list(r.qmin) == [0, -1, 0, -np.inf] and list(r.qmax) == [0, 2, 0.1, np.inf]
    assert np.allclose(r.axis[3], [0, 0, 1])                       # axes are normalised
    assert r.self_collision_edits == [(0, 2, False)]
    d = r.drivers[0]
    assert d.links == [1, 3] and d.scale == [1.0, 2.0] and d.offset == [0.0, 0.1] and (d.qmin, d.qmax) == (-1.0, 2.0)
    assert [w.geoms[g].tris.shape[0] for g in r.link_geom[:3]] == [48, 96, 320] and r.link_geom[3] == -1
    o = OracleWorld(w)
    q = np.array([0.0, 0.5, 0.05, 1.1])
    T = o.fk(q)
    # arm frame: base (0,0,0.1) + (0,0,0.1), yawed 90 deg, pitched 0.5 about its own y; finger 0.5 + 0.05 along the arm's z
    Rz = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]); c, s = math.cos(0.5), math.sin(0.5)
    Ry = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    assert np.allclose(T[1, :9].reshape(3, 3), Rz @ Ry, atol=1e-12) and np.allclose(T[1, 9:], [0, 0, 0.2])
    assert np.allclose(T[2, 9:], np.array([0, 0, 0.2]) + Rz @ Ry @ np.array([0, 0, 0.55]), atol=1e-12)
    # mimic joint = affine driver: CheckJointLimits bounds the driver VALUE, the mean of q1 and (q3 - 0.1) / 2 (Robot.cpp:2166-2187)
    assert o.check_joint_limits(q) and o.check_joint_limits(np.array([0.0, 2.0, 0.05, 4.1]))
    assert not o.check_joint_limits(np.array([0.0, 2.0, 0.05, 6.1]))


def test_urdf_floating_base_gets_six_virtual_links():
    """a root link that is not the world frame makes the robot floating (Robot.cpp:2864-2998)"""
    from klampt_b200.worldspec import JOINT_FLOATING
    text = URDF_ARM.replace('<link name="world"/>', "").replace(
        '<joint name="fix" type="fixed"><parent link="world"/><child link="base"/><origin xyz="0 0 0.1"/></joint>', "")
    w, r = kio.parse_urdf(text)
    assert r.names[:6] == ["base0", "base1", "base2", "base3", "base4", "base"] and list(r.parents[:7]) == [-1, 0, 1, 2, 3, 4, 5]
    assert list(r.linktype[:6]) == [1, 1, 1, 0, 0, 0] and r.axis[3:6].tolist() == [[0, 0, 1], [0, 1, 0], [1, 0, 0]]
    assert r.joint_type[0] == JOINT_FLOATING and r.joint_link[0] == 5 and r.joint_base[0] == -1
    o = OracleWorld(w)                                             # the oracle validates the floating joint's layout
    a = np.zeros(r.L); b = np.zeros(r.L); b[3] = 1.0; b[0] = 0.3
    assert abs(o.cspace_distance(a, b) - math.sqrt(1.0 + 0.09)) < 1e-12
    frozen = kio.parse_urdf(text.replace("<klampt>", '<klampt freeze_root_link="1">'))[1]
    assert list(frozen.joint_type[:6]) == [JOINT_WELD] * 6 and (frozen.qmin[:6] == 0).all()


def test_pcd_stl_and_world_xml(tmp_path):
    """.pcd (ascii / binary), .stl (ascii / binary) and the world file entities the path reads
    (Cpp/docs/Manual-FileTypes.md:47-162)"""
    pcd = b"# .PCD v0.7\nVERSION 0.7\nFIELDS x y z radius\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\nPOINTS 3\nDATA ascii\n0 0 0 0.1\n1 2 3 0.2\nnan 0 0 0.3\n"
    pts, rad = kio.parse_pcd(pcd)
    assert pts.tolist() == [[0, 0, 0], [1, 2, 3]] and np.allclose(rad, [0.1, 0.2])
    rec = np.zeros(2, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4")]); rec["x"] = [1, 2]; rec["z"] = [5, 6]
    binp = b"VERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n" + rec.tobytes()
    pts2, rad2 = kio.parse_pcd(binp)
    assert pts2.tolist() == [[1, 0, 5], [2, 0, 6]] and rad2 is None
    v, t = synth.unit_cube()
    ascii_stl = "solid c\n" + "".join("facet normal 0 0 0\nouter loop\n" + "".join("vertex %g %g %g\n" % tuple(v[i]) for i in tri) + "endloop\nendfacet\n" for tri in t) + "endsolid c\n"
    sv, st = kio.parse_stl(ascii_stl.encode())
    assert st.shape == (12, 3) and np.allclose(sv[st.reshape(-1)], v[t.reshape(-1)])
    recs = np.zeros(12, dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]); recs["v"] = v[t]
    bv, bt = kio.parse_stl(b"\0" * 80 + np.uint32(12).tobytes() + recs.tobytes())
    assert np.allclose(bv[bt.reshape(-1)], v[t.reshape(-1)])
    # world file: terrain (scaled, lifted), rigid object from a point cloud with a pose, robot from a .rob file
    (tmp_path / "cube.off").write_text(CUBE_OFF)
    (tmp_path / "cloud.pcd").write_bytes(pcd)
    spec0 = WorldSpec(); arm = synth.make_planar_nR(spec0, 2)
    (tmp_path / "arm.rob").write_text(kio.rob_text(arm, spec0))
    (tmp_path / "w.xml").write_text("""<?xml version="1.0"?>
<world>
  <terrain file="cube.off" scale="4 4 0.1" translation="-2 -2 -0.1" margin="0.01"><display color="1 0 0"/></terrain>
  <rigidObject name="blob" position="1 0 0.5" rotateZ="1.5707963267948966">
    <geometry file="cloud.pcd" scale="0.5" margin="0.02"/><physics mass="1"/>
  </rigidObject>
  <robot name="arm" file="arm.rob"/>
  <simulation><globals maxContacts="20"/></simulation>
</world>""")
    w = kio.load_world_xml(str(tmp_path / "w.xml"))
    assert len(w.terrains) == 1 and len(w.objects) == 1 and w.robot.L == 2
    g = w.geoms[w.terrains[0]]
    assert np.allclose(g.verts.min(0), [-2, -2, -0.1]) and np.allclose(g.verts.max(0), [2, 2, 0.0]) and g.margin == 0.01
    go, T = w.objects[0]
    assert w.geoms[go].kind == "cloud" and np.allclose(w.geoms[go].points, [[0, 0, 0], [0.5, 1.0, 1.5]]) and np.allclose(w.geoms[go].radius, [0.05, 0.1])
    assert np.allclose(T[:9].reshape(3, 3), [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-12) and np.allclose(T[9:], [1, 0, 0.5])
    o = OracleWorld(w)                                             # the whole description loads into the checker
    assert o.num_ids() == 1 + 1 + 1 + 2


def test_rob_floating_joint_base_round_trip():
    w = synth.world_floating(n_obstacles=2)
    text = kio.rob_text(w.robot, w)
    assert "joint floating 5 -1" in text and "joint ballandsocket 9 6" in text
    w2, r2 = kio.parse_rob(text)
    assert list(r2.joint_base) == list(w.robot.joint_base) and list(r2.joint_type) == list(w.robot.joint_type)


def test_world_xml_places_the_robot_and_reads_transforms_like_the_reference(tmp_path):
    """XmlRobot::GetRobot pre-multiplies the <robot> element's transform into T0_Parent of every root link (Cpp/IO/XmlWorld.cpp:251-257);
    ReadTransform (:97-160) takes translation else position, and applies rotateRPY (set) -> rotateMoment (replace) -> rotateX -> rotateY
    -> rotateZ in that fixed order whatever the attribute order; <geometry> has its own transform (:292-294)"""
    (tmp_path / "cube.off").write_text(CUBE_OFF)
    spec0 = WorldSpec(); arm = synth.make_planar_nR(spec0, 2)
    (tmp_path / "arm.rob").write_text(kio.rob_text(arm, spec0))
    (tmp_path / "w.xml").write_text("""<?xml version="1.0"?>
<world>
  <robot name="arm" file="arm.rob" position="1 2 0.5" rotateZ="1.5707963267948966"/>
  <rigidObject name="box" rotateZ="0.3" rotateX="0.2" position="0 0 1">
    <geometry file="cube.off" scale="2 1 1" rotateZ="1.5707963267948966" translation="0 0 3"/>
  </rigidObject>
</world>""")
    w = kio.load_world_xml(str(tmp_path / "w.xml"))
    base = kio.parse_rob(kio.rob_text(arm, spec0))[1]
    Rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    R0, t0 = base.T0[0, :9].reshape(3, 3), base.T0[0, 9:12]
    assert np.allclose(w.robot.T0[0, :9].reshape(3, 3), Rz @ R0, atol=1e-12) and np.allclose(w.robot.T0[0, 9:12], Rz @ t0 + [1, 2, 0.5], atol=1e-12)
    assert np.allclose(w.robot.T0[1], base.T0[1])                      # only root links move
    # the displaced robot really is somewhere else for the feasibility path: link 0's world transform carries the offset
    o = OracleWorld(w)
    assert np.allclose(o.fk(np.zeros(2))[0, 9:12], Rz @ t0 + [1, 2, 0.5], atol=1e-12)
    # fixed X-then-Z order although the file writes Z first:  R = Rz(0.3) Rx(0.2)
    from klampt_b200 import so3
    go, T = w.objects[0]
    want = so3.matrix(so3.from_axis_angle(([0, 0, 1], 0.3))) @ so3.matrix(so3.from_axis_angle(([1, 0, 0], 0.2)))
    assert np.allclose(T[:9].reshape(3, 3), want, atol=1e-12) and np.allclose(T[9:], [0, 0, 1])
    # geometry: scale (2,1,1), then quarter turn about z, then lift by 3: the unit cube spans x in [-1,0], y in [0,2], z in [3,4]
    v = w.geoms[go].verts
    assert np.allclose(v.min(0), [-1, 0, 3], atol=1e-12) and np.allclose(v.max(0), [0, 2, 4], atol=1e-12)


def test_rob_single_value_broadcast_and_selfcollision_names_of_mounted_links(tmp_path):
    """geomscale / geommargin with one value apply to every link (Robot.cpp:882-887,1024-1051); self-collision pairs naming links
    of a sub-chain mounted further down are resolved after the mounts (Robot.cpp:1297-1313,1344-1380)"""
    (tmp_path / "tri.off").write_text("OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    (tmp_path / "hand.rob").write_text('links palm finger\nparents -1 0\ntparent 1 0 0 0 1 0 0 0 1 0 0 0  1 0 0 0 1 0 0 0 1 0.1 0 0\n'
                                       'qmin -1 -1\nqmax 1 1\ngeometry "tri.off" "tri.off"\n')
    text = ('links a b\nparents -1 0\ntparent 1 0 0 0 1 0 0 0 1 0 0 0  1 0 0 0 1 0 0 0 1 1 0 0\nqmin -1 -1\nqmax 1 1\n'
            'geometry "tri.off" "tri.off"\ngeomscale 2\ngeommargin 0.01\nnoselfcollision a hand:finger\nmount 1 "hand.rob" as "hand"\n')
    world, r = kio.parse_rob(text, basedir=str(tmp_path))
    assert r.L == 4 and list(r.names) == ["a", "b", "hand:palm", "hand:finger"]
    assert (0, 3, False) in r.self_collision_edits
    for i in (0, 1):
        g = world.geoms[r.link_geom[i]]
        assert np.isclose(g.verts.max(), 2.0) and np.isclose(g.margin, 0.01)
    with pytest.raises(ValueError):
        kio.parse_rob(text.replace("hand:finger", "hand:thumb"), basedir=str(tmp_path))
    with pytest.raises(ValueError):
        kio.parse_rob("links a b\nparents -1 0\nnoselfcollision a zz\n")
