"""Ingestion of the reference's data formats into the engine's plain world description, so the
engine can be fed without a Klamp't install:

  * triangle meshes: OFF (the format of the reference's ``tests/objects/cube.off``) and Wavefront OBJ
    (``tests/objects/block.obj``), read and written;
  * robots: the kinematic / geometric / joint subset of the ``.rob`` format documented in
    Cpp/docs/Manual-FileTypes.md:163-236 and parsed by Cpp/Modeling/Robot.cpp:216-1383 that the feasibility path reads:
    ``links parents jointtype tparent axis qmin/qmax(deg) q translation rotation scale geometry geomscale geommargin
    noselfcollision selfcollision joint driver``.  Inline geometry strings (``"{TriangleMesh\\nOFF ...}"`` as written by
    Python/klampt/model/create/planar_robot.py:20-70) are understood.  Dynamic items (mass, inertia, torque limits,
    servo gains ...) are parsed over and ignored.  ``mount``, D-H parameters and URDF are not supported.

Root links get ``rotation`` / ``translation`` pre-multiplied into ``T0_Parent`` as Robot.cpp:971-975 does.
"""
from __future__ import annotations

import math
import os
import shlex
from typing import Dict, List, Optional, Tuple

import numpy as np

from .worldspec import (GeomSpec, RobotSpec, WorldSpec, DriverSpec, REVOLUTE, PRISMATIC, JOINT_WELD, JOINT_NORMAL, JOINT_SPIN,
                        JOINT_FLOATING, JOINT_FLOATINGPLANAR, JOINT_BALLANDSOCKET, IDENTITY12)

_JOINT_TYPES = {"weld": JOINT_WELD, "normal": JOINT_NORMAL, "spin": JOINT_SPIN, "floating": JOINT_FLOATING,
                "floatingplanar": JOINT_FLOATINGPLANAR, "ballandsocket": JOINT_BALLANDSOCKET}


# --------------------------------------------------------------------------------------- meshes
def _off_tokens(text: str) -> List[str]:
    toks = []
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if line:
            toks += line.split()
    return toks


def parse_off(text: str) -> Tuple[np.ndarray, np.ndarray]:
    """OFF text -> (vertices (n,3) f64, triangles (m,3) i32); polygons are fan-triangulated"""
    toks = _off_tokens(text)
    if not toks or not toks[0].upper().startswith("OFF"):
        raise ValueError("not an OFF file")
    toks = toks[1:] if toks[0].upper() == "OFF" else [toks[0][3:]] + toks[1:]
    nv, nf = int(toks[0]), int(toks[1])
    pos = 3
    verts = np.array(toks[pos:pos + 3 * nv], dtype=np.float64).reshape(nv, 3)
    pos += 3 * nv
    tris = []
    for _ in range(nf):
        k = int(toks[pos])
        idx = [int(t) for t in toks[pos + 1:pos + 1 + k]]
        pos += 1 + k
        for j in range(1, k - 1):
            tris.append([idx[0], idx[j], idx[j + 1]])
    tris = np.array(tris, dtype=np.int32).reshape(-1, 3)
    if len(tris) and (tris.min() < 0 or tris.max() >= nv):
        raise ValueError("OFF face references a vertex out of range")
    return verts, tris


def parse_obj(text: str) -> Tuple[np.ndarray, np.ndarray]:
    verts, tris = [], []
    for line in text.splitlines():
        p = line.split("#", 1)[0].split()
        if not p:
            continue
        if p[0] == "v":
            verts.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            idx = [int(tok.split("/")[0]) for tok in p[1:]]
            idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
            for j in range(1, len(idx) - 1):
                tris.append([idx[0], idx[j], idx[j + 1]])
    return np.array(verts, dtype=np.float64).reshape(-1, 3), np.array(tris, dtype=np.int32).reshape(-1, 3)


def load_mesh(path: str) -> Tuple[np.ndarray, np.ndarray]:
    text = open(path).read()
    ext = os.path.splitext(path)[1].lower()
    if ext == ".off":
        return parse_off(text)
    if ext == ".obj":
        return parse_obj(text)
    raise ValueError("unsupported mesh format %r (OFF and OBJ are)" % ext)


def off_text(verts, tris) -> str:
    out = ["OFF", "%d %d 0" % (len(verts), len(tris))]
    out += ["%.17g %.17g %.17g" % tuple(v) for v in np.asarray(verts, dtype=np.float64)]
    out += ["3 %d %d %d" % tuple(t) for t in np.asarray(tris)]
    return "\n".join(out) + "\n"


def save_off(path: str, verts, tris):
    with open(path, "w") as f:
        f.write(off_text(verts, tris))


# --------------------------------------------------------------------------------------- .rob
def _logical_lines(text: str) -> List[str]:
    """comments (#...) stripped outside quotes; a backslash that ends a line outside quotes continues it"""
    lines, cur = [], ""
    for raw in text.splitlines():
        out, inq = [], False
        for ch in raw:
            if ch == '"':
                inq = not inq
            if ch == "#" and not inq:
                break
            out.append(ch)
        s = "".join(out).rstrip()
        if s.endswith("\\") and not inq:
            cur += s[:-1] + " "
            continue
        cur += s
        if cur.strip():
            lines.append(cur.strip())
        cur = ""
    if cur.strip():
        lines.append(cur.strip())
    return lines


def _floats(items, scale=1.0):
    return [float("inf") if t.lower() == "inf" else (float("-inf") if t.lower() == "-inf" else float(t)) * scale for t in items]


def _geometry_from_token(tok: str, basedir: str):
    if tok == "" or tok.lower() == "none":
        return None
    if tok.startswith("{"):                                   # inline: {TriangleMesh\nOFF ...}
        body = tok.strip("{}").replace("\\n", "\n")
        kind, _, rest = body.partition("\n")
        if kind.strip() != "TriangleMesh":
            raise ValueError("inline geometry type %r is not supported" % kind.strip())
        return parse_off(rest)
    return load_mesh(tok if os.path.isabs(tok) else os.path.join(basedir, tok))


def load_rob(path: str, world: Optional[WorldSpec] = None) -> Tuple[WorldSpec, RobotSpec]:
    """Parses a .rob file into a RobotSpec whose link geometries are added to ``world`` (created if None)."""
    return parse_rob(open(path).read(), os.path.dirname(os.path.abspath(path)), world)


def parse_rob(text: str, basedir: str = ".", world: Optional[WorldSpec] = None) -> Tuple[WorldSpec, RobotSpec]:
    world = WorldSpec() if world is None else world
    items: Dict[str, List[str]] = {}
    joints: List[Tuple[int, int, int]] = []
    drivers: List[List[str]] = []
    selfcol, noselfcol = [], []
    for line in _logical_lines(text):
        lex = shlex.shlex(line, posix=True)
        lex.whitespace_split = True
        lex.escape = ""                                       # keep the \n escapes of inline geometry strings
        lex.commenters = ""
        toks = list(lex)
        key, args = toks[0].lower(), toks[1:]
        if key == "joint":
            t = _JOINT_TYPES.get(args[0].lower())
            if t is None:
                raise ValueError("unknown joint type %r" % args[0])
            joints.append((t, int(args[1]), int(args[2]) if len(args) > 2 else -1))
        elif key == "driver":
            drivers.append(args)
        elif key == "selfcollision":
            selfcol += args
        elif key == "noselfcollision":
            noselfcol += args
        elif key in ("mount", "alpha", "alphadeg", "a", "d", "theta", "thetadeg"):
            raise NotImplementedError(".rob item %r is not supported by this loader" % key)
        else:
            items[key] = args
    if "parents" not in items:
        raise ValueError(".rob file has no 'parents' line")
    parents = np.array([int(x) for x in items["parents"]], dtype=np.int32)
    L = len(parents)
    names = items.get("links", ["link%d" % i for i in range(L)])

    def per_link(key, default, width=1, scale=1.0):
        if key not in items:
            return np.tile(np.asarray(default, dtype=np.float64), (L, 1)) if width > 1 else np.full(L, default, dtype=np.float64)
        v = np.array(_floats(items[key], scale), dtype=np.float64)
        if v.size != L * width:
            raise ValueError("'%s' needs %d values, got %d" % (key, L * width, v.size))
        return v.reshape(L, width) if width > 1 else v

    if "tparent" not in items:
        raise ValueError(".rob file has no 'tparent' line (D-H parameters are not supported)")
    T0 = per_link("tparent", None, 12)
    axis = per_link("axis", [0.0, 0.0, 1.0], 3)
    deg = math.pi / 180.0
    qmin = per_link("qmindeg", 0, 1, deg) if "qmindeg" in items else per_link("qmin", -np.inf)
    qmax = per_link("qmaxdeg", 0, 1, deg) if "qmaxdeg" in items else per_link("qmax", np.inf)
    linktype = np.array([PRISMATIC if t.lower().startswith("p") else REVOLUTE for t in items.get("jointtype", ["r"] * L)], dtype=np.uint8)
    scale = float(items["scale"][0]) if "scale" in items else 1.0
    if scale != 1.0:
        T0[:, 9:12] *= scale
    # base transform of root links (Robot.cpp:971-975)
    Rb = np.array(_floats(items["rotation"]), dtype=np.float64).reshape(3, 3) if "rotation" in items else np.eye(3)
    tb = np.array(_floats(items["translation"]), dtype=np.float64) if "translation" in items else np.zeros(3)
    for i in range(L):
        if parents[i] < 0:
            R, t = T0[i, :9].reshape(3, 3), T0[i, 9:12]
            T0[i, :9] = (Rb @ R).reshape(-1)
            T0[i, 9:12] = Rb @ t + tb
    # geometry
    gscale = per_link("geomscale", 1.0)
    gmargin = per_link("geommargin", 0.0)
    link_geom = [-1] * L
    cache: Dict[Tuple[str, float], int] = {}
    for i, tok in enumerate(items.get("geometry", [""] * L)):
        mesh = _geometry_from_token(tok, basedir)
        if mesh is None or len(mesh[1]) == 0:
            continue
        v, t = mesh
        link_geom[i] = world.add_geom(GeomSpec.mesh(v * gscale[i] * scale, t, margin=gmargin[i]))
    # joints: default one Normal joint per link
    if joints:
        jt = np.array([j[0] for j in joints], dtype=np.uint8)
        jl = np.array([j[1] for j in joints], dtype=np.int32)
    else:
        jt, jl = np.full(L, JOINT_NORMAL, dtype=np.uint8), np.arange(L, dtype=np.int32)

    def link_index(tok: str) -> int:
        try:
            return int(tok)
        except ValueError:
            return names.index(tok)

    edits = []
    for lst, en in ((selfcol, True), (noselfcol, False)):
        if len(lst) % 2:
            raise ValueError("self-collision lists hold link PAIRS")
        if en and lst:
            # an explicit selfcollision list replaces the default set (Robot.cpp:1277-1296): start from nothing
            edits += [(i, j, False) for i in range(L) for j in range(i + 1, L)]
        for a, b in zip(lst[0::2], lst[1::2]):
            i, j = link_index(a), link_index(b)
            if i == j:
                raise ValueError("Invalid self collision pair %s, %s" % (a, b))
            edits.append((min(i, j), max(i, j), en))
    drv = []
    for d in drivers:
        kind = d[0].lower()
        if kind == "normal":
            k = int(d[1])
            drv.append(DriverSpec([k], [1.0], [0.0], float(qmin[k]), float(qmax[k])))
        elif kind == "affine":
            n = int(d[1])
            links = [int(x) for x in d[2:2 + n]]
            sc = _floats(d[2 + n:2 + 2 * n])
            of = _floats(d[2 + 2 * n:2 + 3 * n])
            rest = _floats(d[2 + 3 * n:])
            drv.append(DriverSpec(links, sc, of, rest[0] if len(rest) > 0 else -np.inf, rest[1] if len(rest) > 1 else np.inf))
    spec = RobotSpec(parents=parents, linktype=linktype, axis=axis, T0=T0, qmin=qmin, qmax=qmax, link_geom=link_geom, joint_type=jt, joint_link=jl,
                     drivers=drv, self_collision_edits=edits, names=list(names))
    world.robot = spec
    return world, spec


def rob_text(spec: RobotSpec, world: WorldSpec) -> str:
    """Writes the kinematic / geometric subset back out with inline OFF geometry (round-trips through parse_rob)."""
    L = spec.L
    names = spec.names or ["link%d" % i for i in range(L)]
    fmt = lambda a: " ".join("inf" if np.isposinf(x) else ("-inf" if np.isneginf(x) else "%.17g" % x) for x in np.asarray(a, dtype=np.float64).reshape(-1))
    out = ["links " + " ".join('"%s"' % n for n in names),
           "parents " + " ".join(str(int(p)) for p in spec.parents),
           "jointtype " + " ".join("p" if t == PRISMATIC else "r" for t in spec.linktype),
           "tparent " + " \\\n".join(fmt(T) for T in spec.T0),
           "axis " + fmt(spec.axis), "qmin " + fmt(spec.qmin), "qmax " + fmt(spec.qmax)]
    geo, margins = [], []
    for gi in spec.link_geom:
        if gi < 0:
            geo.append('""')
            margins.append(0.0)
        else:
            g = world.geoms[gi]
            geo.append('"{TriangleMesh\\n' + off_text(g.verts, g.tris).replace("\n", "\\n") + '}"')
            margins.append(g.margin)
    out.append("geometry " + " ".join(geo))
    out.append("geommargin " + fmt(margins))
    inv = {v: k for k, v in _JOINT_TYPES.items()}
    if spec.joint_type is not None:
        for t, k in zip(spec.joint_type, spec.joint_link):
            out.append("joint %s %d" % (inv[int(t)], int(k)))
    dis = [(i, j) for (i, j, en) in spec.self_collision_edits if not en]
    if dis:
        out.append("noselfcollision " + " ".join("%d %d" % p for p in dis))
    return "\n".join(out) + "\n"
