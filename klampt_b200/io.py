"""Ingestion of the reference's data formats into the engine's plain world description, so the
engine can be fed without a Klamp't install:

  * triangle meshes: OFF (the format of the reference's ``tests/objects/cube.off``), Wavefront OBJ
    (``tests/objects/block.obj``) and STL (ascii / binary); point clouds: PCD (ascii / uncompressed binary, optional
    ``radius`` field);
  * robots: the kinematic / geometric / joint subset of the ``.rob`` format documented in
    Cpp/docs/Manual-FileTypes.md:163-236 and parsed by Cpp/Modeling/Robot.cpp:216-1383 that the feasibility path reads:
    ``links parents jointtype tparent axis qmin/qmax(deg) q translation rotation scale geometry geomscale geommargin
    noselfcollision selfcollision joint driver``.  Inline geometry strings (``"{TriangleMesh\\nOFF ...}"`` as written by
    Python/klampt/model/create/planar_robot.py:20-70) are understood.  Dynamic items (mass, inertia, torque limits,
    servo gains ...) are parsed over and ignored.  ``mount`` (geometry files and .rob / .urdf sub-chains) is supported, D-H
    parameters are not;
  * URDF with the ``<klampt>`` element (Manual-FileTypes.md:240-274), built the way RobotModel::LoadURDF builds its links
    (Robot.cpp:2566-3300): fixed or floating base, revolute / continuous / prismatic / fixed joints, mimic joints as affine
    drivers, box / cylinder / sphere / mesh collision geometry;
  * world files (Manual-FileTypes.md:47-162): terrains, rigid objects and the robot;
  * configurations and transforms: ``.config`` / ``.configs`` / ``.xform`` text as Python/klampt/io/loader.py writes it.

Root links get ``rotation`` / ``translation`` pre-multiplied into ``T0_Parent`` as Robot.cpp:971-975 does.
"""
from __future__ import annotations

import math
import os
import re
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


def parse_stl(data: bytes) -> Tuple[np.ndarray, np.ndarray]:
    """ASCII or binary STL -> (vertices, triangles); vertices are not merged (every triangle keeps its own three)."""
    head = data[:512].lstrip().lower()
    if head.startswith(b"solid") and b"facet" in data[:4096].lower():
        toks = data.decode("ascii", "replace").split()
        v = [float(toks[i + k]) for i, t in enumerate(toks) if t.lower() == "vertex" for k in (1, 2, 3)]
        verts = np.array(v, dtype=np.float64).reshape(-1, 3)
    else:
        n = int(np.frombuffer(data[80:84], dtype="<u4")[0])
        rec = np.frombuffer(data[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        verts = rec["v"].reshape(-1, 3).astype(np.float64)
    return verts, np.arange(len(verts), dtype=np.int32).reshape(-1, 3)


def parse_pcd(data: bytes) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Point Cloud Data file (ascii or uncompressed binary) -> (points (n,3), radius (n,) or None).  Only the x, y, z and
    radius fields are read (the `radius` property makes every point a sphere, Cpp/docs/Manual-Geometry.md:25); rows with a
    non-finite coordinate are dropped."""
    hdr: Dict[str, List[str]] = {}
    pos = 0
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, *vals = line.split()
        hdr[key.upper()] = vals
        if key.upper() == "DATA":
            break
    fields = hdr["FIELDS"]
    sizes = [int(x) for x in hdr.get("SIZE", ["4"] * len(fields))]
    types = hdr.get("TYPE", ["F"] * len(fields))
    counts = [int(x) for x in hdr.get("COUNT", ["1"] * len(fields))]
    npts = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr.get("HEIGHT", ["1"])[0])
    mode = hdr["DATA"][0].lower()
    cols = {}
    if mode == "ascii":
        rows = np.array(data[pos:].decode("ascii", "replace").split(), dtype=np.float64).reshape(npts, -1)
        c = 0
        for f, k in zip(fields, counts):
            cols[f] = rows[:, c]
            c += k
    elif mode == "binary":
        dt = []
        for f, sz, ty, k in zip(fields, sizes, types, counts):
            code = {"F": "f", "I": "i", "U": "u"}[ty.upper()] + str(sz)
            dt.append((f, "<" + code, (k,)) if k > 1 else (f, "<" + code))
        rec = np.frombuffer(data[pos:pos + np.dtype(dt).itemsize * npts], dtype=np.dtype(dt))
        for f in fields:
            cols[f] = np.asarray(rec[f], dtype=np.float64).reshape(npts, -1)[:, 0]
    else:
        raise ValueError("PCD DATA mode %r is not supported (ascii and binary are)" % mode)
    pts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
    ok = np.isfinite(pts).all(axis=1)
    rad = np.asarray(cols["radius"], dtype=np.float64)[ok] if "radius" in cols else None
    return np.ascontiguousarray(pts[ok]), rad


def load_mesh(path: str) -> Tuple[np.ndarray, np.ndarray]:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".stl":
        return parse_stl(open(path, "rb").read())
    text = open(path).read()
    if ext == ".off":
        return parse_off(text)
    if ext == ".obj":
        return parse_obj(text)
    raise ValueError("unsupported mesh format %r (OFF, OBJ and STL are)" % ext)


def load_geometry(path: str, scale=1.0, translate=(0.0, 0.0, 0.0), margin: float = 0.0) -> GeomSpec:
    """any geometry file the path reads (.off / .obj / .stl mesh, .pcd point cloud) as a GeomSpec; `scale` may be a 3-vector"""
    sc = np.broadcast_to(np.asarray(scale, dtype=np.float64), (3,))
    tr = np.asarray(translate, dtype=np.float64)
    if os.path.splitext(path)[1].lower() == ".pcd":
        pts, rad = parse_pcd(open(path, "rb").read())
        return GeomSpec.cloud(pts * sc + tr, None if rad is None else rad * float(sc.max()), margin)
    v, t = load_mesh(path)
    return GeomSpec.mesh(v * sc + tr, t, margin)


def off_text(verts, tris) -> str:
    out = ["OFF", "%d %d 0" % (len(verts), len(tris))]
    out += ["%.17g %.17g %.17g" % tuple(v) for v in np.asarray(verts, dtype=np.float64)]
    out += ["3 %d %d %d" % tuple(t) for t in np.asarray(tris)]
    return "\n".join(out) + "\n"


def save_off(path: str, verts, tris):
    with open(path, "w") as f:
        f.write(off_text(verts, tris))


# --------------------------------------------------------------------------------------- configurations / transforms
# The reference's resource text formats (Python/klampt/io/loader.py:161-198,226-250; the same text the C++ bindings write):
#   .config   "n<TAB>v1 v2 ... vn"        one configuration
#   .configs  one such record after the other (any whitespace)  -> the N x L batch the engine takes
#   .xform    "r11 r12 r13 r21 ... r33 t1 t2 t3"  row-major R then t = the 12-vector of the C ABI
def write_config(q) -> str:
    q = [float(v) for v in q]
    return str(len(q)) + "\t" + " ".join(repr(v) for v in q)


def read_config(text: str) -> np.ndarray:
    items = text.split()
    if len(items) == 0:
        raise ValueError("Empty text")
    if int(items[0]) + 1 != len(items):
        raise ValueError("Invalid number of items")
    return np.array([float(v) for v in items[1:]], dtype=np.float64)


def write_configs(Q) -> str:
    return "\n".join(write_config(q) for q in np.asarray(Q, dtype=np.float64).reshape(len(Q), -1))


def read_configs(text: str) -> np.ndarray:
    """all records of a .configs text as one (N, L) array; records of different lengths are an error (a batch is rectangular)"""
    items, rows, pos = text.split(), [], 0
    while pos < len(items):
        n = int(items[pos])
        if pos + 1 + n > len(items):
            raise ValueError("Invalid number of items")
        rows.append([float(v) for v in items[pos + 1:pos + 1 + n]])
        pos += 1 + n
    if rows and any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("configurations of different lengths cannot form a batch")
    return np.array(rows, dtype=np.float64).reshape(len(rows), len(rows[0]) if rows else 0)


def write_xform(T12) -> str:
    T = np.asarray(T12, dtype=np.float64).reshape(12)
    return "\t".join(" ".join(repr(float(v)) for v in T[3 * i:3 * i + 3]) for i in range(3)) + "\t" + " ".join(repr(float(v)) for v in T[9:])


def read_xform(text: str) -> np.ndarray:
    items = text.split()
    if len(items) != 12:
        raise ValueError("Invalid element of SE3, must have 12 elements")
    return np.array([float(v) for v in items], dtype=np.float64)


def load_configs(path: str) -> np.ndarray:
    return read_configs(open(path).read())


def save_configs(path: str, Q) -> None:
    with open(path, "w") as f:
        f.write(write_configs(Q) + "\n")


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


_NUM_PREFIX = re.compile(r"[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")


def _floats(items, scale=1.0):
    """numbers of one .rob line with the semantics of the reference's ``while (ss >> ftemp)`` loops (Cpp/Modeling/Robot.cpp:409-431):
    reading stops at the first token that is not a number, and a token that only STARTS with a number still yields that number.
    The reference's own planar-robot generator depends on it: it ends its TParent line with a literal backslash-n
    (model/create/planar_robot.py:44), which glues the following ``axis`` line onto it -- the reader takes the last translation
    entry from the token ``0\\naxis`` and drops the rest, axes included (tests/golden/ref_planar_3R.rob)."""
    out = []
    for t in items:
        low = t.lower()
        if low in ("inf", "+inf", "-inf"):
            out.append(float(low))
            continue
        m = _NUM_PREFIX.match(t)
        if m is None:
            break
        out.append(float(m.group(0)) * scale)
        if m.end() != len(t):
            break
    return out


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
    mounts: List[List[str]] = []
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
        elif key == "mount":
            mounts.append(args)
        elif key in ("alpha", "alphadeg", "a", "d", "theta", "thetadeg"):
            raise NotImplementedError(".rob item %r is not supported by this loader" % key)
        else:
            items[key] = args
    if "parents" in items:
        parents = np.array([int(x) for x in items["parents"]], dtype=np.int32)
    elif "tparent" in items:                 # no parents line: a serial chain, parents[i] = i - 1 (Robot.cpp:899-902)
        parents = np.arange(-1, len(_floats(items["tparent"])) // 12 - 1, dtype=np.int32)
    else:
        raise ValueError(".rob file has neither a 'parents' nor a 'tparent' line")
    L = len(parents)
    names = items.get("links", ["Link_%d" % i for i in range(L)])          # the reference's default names (Robot.cpp:907-914)

    def per_link(key, default, width=1, scale=1.0, broadcast=False):
        if key not in items:
            return np.tile(np.asarray(default, dtype=np.float64), (L, 1)) if width > 1 else np.full(L, default, dtype=np.float64)
        v = np.array(_floats(items[key], scale), dtype=np.float64)
        if broadcast and v.size == 1 and width == 1:       # one value for every link (geomscale / geommargin, Robot.cpp:882-887,1024-1051)
            return np.full(L, v[0], dtype=np.float64)
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
    gscale = per_link("geomscale", 1.0, broadcast=True)
    gmargin = per_link("geommargin", 0.0, broadcast=True)
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
        # "joint floating <link> <base>": the third field is RobotModelJoint::baseIndex (Robot.cpp:1152-1180); single-link joints hang
        # from their link's parent
        jb = np.array([j[2] if j[0] in (JOINT_FLOATING, JOINT_FLOATINGPLANAR, JOINT_BALLANDSOCKET) else int(parents[j[1]]) for j in joints], dtype=np.int32)
    else:
        jt, jl, jb = np.full(L, JOINT_NORMAL, dtype=np.uint8), np.arange(L, dtype=np.int32), None

    def link_index(tok: str) -> int:
        try:
            return int(tok)
        except ValueError:
            return names.index(tok)

    edits = []
    residual = []        # pairs that name links of a sub-chain mounted further down: resolved after the mounts (Robot.cpp:1297-1313,1344-1380)
    for lst, en in ((selfcol, True), (noselfcol, False)):
        if len(lst) % 2:
            raise ValueError("self-collision lists hold link PAIRS")
        if en and lst:
            # an explicit selfcollision list replaces the default set (Robot.cpp:1277-1296): start from nothing
            edits += [(i, j, False) for i in range(L) for j in range(i + 1, L)]
        for a, b in zip(lst[0::2], lst[1::2]):
            try:
                i, j = link_index(a), link_index(b)
            except ValueError:
                if not mounts:
                    raise ValueError("self-collision pair names an unknown link: %s, %s" % (a, b))
                residual.append((a, b, en))
                continue
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
                     joint_base=jb, drivers=drv, self_collision_edits=edits, names=list(names))
    for m in mounts:
        _mount(world, spec, m, basedir, link_index)
    for a, b, en in residual:
        def late(tok):
            try:
                return int(tok)
            except ValueError:
                return list(spec.names).index(tok)
        try:
            i, j = late(a), late(b)
        except ValueError:
            raise ValueError("self-collision pair names an unknown link (also after the mounts): %s, %s" % (a, b))
        if i == j:
            raise ValueError("Invalid self collision pair %s, %s" % (a, b))
        spec.self_collision_edits.append((min(i, j), max(i, j), en))
    world.robot = spec
    return world, spec


def _mount(world: WorldSpec, spec: RobotSpec, args: List[str], basedir: str, link_index) -> None:
    """``mount <link> "<file>" [R t] [as "<prefix>"]`` (Robot.cpp:648-690): a geometry file is merged into the link's geometry under the
    transform (Robot.cpp:1213-1245, RobotModel::Mount :1860-1893); a .rob / .urdf file is appended as a sub-chain whose roots hang from
    the link with `T * T0_Parent`, with links, joints and drivers renumbered and link names prefixed "prefix:" (:1895-2007).  Self
    collisions between the old links and the new ones follow the default rule (all pairs but parent / child), which is what
    InitSelfCollisionPair over all (i, j + norig) minus the mount link amounts to."""
    if len(args) < 2:
        raise ValueError("mount needs a link and a file")
    link = -1 if args[0] == "-1" else link_index(args[0])
    fn = args[1]
    rest = args[2:]
    T = np.array(IDENTITY12, dtype=np.float64)
    nums = _floats(rest[:12])
    if len(nums) == 12:
        T, rest = np.array(nums), rest[12:]
    prefix = rest[1] if len(rest) >= 2 and rest[0].lower() == "as" else None
    path = fn if os.path.isabs(fn) else os.path.join(basedir, fn)
    R, t = T[:9].reshape(3, 3), T[9:12]
    if os.path.splitext(fn)[1].lower() not in (".rob", ".urdf"):
        v, tr = load_mesh(path)
        v = v @ R.T + t
        if link < 0:
            raise ValueError("a geometry can only be mounted on a link")
        if spec.link_geom[link] >= 0 and world.geoms[spec.link_geom[link]].kind == "mesh":
            g = world.geoms[spec.link_geom[link]]
            v, tr = np.vstack([g.verts, v]), np.vstack([g.tris, tr + len(g.verts)]).astype(np.int32)
            spec.link_geom[link] = world.add_geom(GeomSpec.mesh(v, tr, margin=g.margin))
        else:
            spec.link_geom[link] = world.add_geom(GeomSpec.mesh(v, tr))
        return
    keep = world.robot
    _, sub = (load_urdf(path, world) if path.lower().endswith(".urdf") else load_rob(path, world))
    world.robot = keep
    n0 = spec.L
    T0 = sub.T0.copy()
    par = sub.parents.copy()
    for i in range(sub.L):
        if sub.parents[i] < 0:
            Rs, ts = T0[i, :9].reshape(3, 3), T0[i, 9:12]
            T0[i, :9], T0[i, 9:12] = (R @ Rs).reshape(-1), R @ ts + t
            par[i] = link
        else:
            par[i] += n0
    spec.parents = np.concatenate([spec.parents, par]).astype(np.int32)
    spec.linktype = np.concatenate([spec.linktype, sub.linktype]).astype(np.uint8)
    spec.axis, spec.T0 = np.vstack([spec.axis, sub.axis]), np.vstack([spec.T0, T0])
    spec.qmin, spec.qmax = np.concatenate([spec.qmin, sub.qmin]), np.concatenate([spec.qmax, sub.qmax])
    spec.link_geom = list(spec.link_geom) + list(sub.link_geom)
    sub_names = sub.names or ["Link_%d" % i for i in range(sub.L)]
    spec.names = list(spec.names or ["Link_%d" % i for i in range(n0)]) + [(prefix + ":" + s) if prefix else s for s in sub_names]
    sjt = sub.joint_type if sub.joint_type is not None else np.full(sub.L, JOINT_NORMAL, dtype=np.uint8)
    sjl = sub.joint_link if sub.joint_link is not None else np.arange(sub.L, dtype=np.int32)
    sjb = sub.joint_base if sub.joint_base is not None else sub.parents[sjl]
    jb0 = spec.joint_base if spec.joint_base is not None else spec.parents[:n0][spec.joint_link]
    spec.joint_type = np.concatenate([spec.joint_type, sjt]).astype(np.uint8)
    spec.joint_link = np.concatenate([spec.joint_link, sjl + n0]).astype(np.int32)
    spec.joint_base = np.concatenate([jb0, np.where(np.asarray(sjb) < 0, link, np.asarray(sjb) + n0)]).astype(np.int32)
    spec.drivers = list(spec.drivers) + [DriverSpec([k + n0 for k in d.links], list(d.scale), list(d.offset), d.qmin, d.qmax) for d in sub.drivers]
    spec.self_collision_edits = list(spec.self_collision_edits) + [(i + n0, j + n0, en) for i, j, en in sub.self_collision_edits]


def rob_text(spec: RobotSpec, world: WorldSpec) -> str:
    """Writes the kinematic / geometric subset back out with inline OFF geometry (round-trips through parse_rob)."""
    L = spec.L
    names = spec.names or ["Link_%d" % i for i in range(L)]
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
        for n, (t, k) in enumerate(zip(spec.joint_type, spec.joint_link)):
            if int(t) in (JOINT_FLOATING, JOINT_FLOATINGPLANAR, JOINT_BALLANDSOCKET) and spec.joint_base is not None:
                out.append("joint %s %d %d" % (inv[int(t)], int(k), int(spec.joint_base[n])))
            else:
                out.append("joint %s %d" % (inv[int(t)], int(k)))
    dis = [(i, j) for (i, j, en) in spec.self_collision_edits if not en]
    if dis:
        out.append("noselfcollision " + " ".join("%d %d" % p for p in dis))
    return "\n".join(out) + "\n"


# ============================================================================================== URDF
def _rpy_matrix(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)"""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(el) -> Tuple[np.ndarray, np.ndarray]:
    if el is None:
        return np.eye(3), np.zeros(3)
    xyz = np.array([float(x) for x in el.get("xyz", "0 0 0").split()], dtype=np.float64)
    rpy = [float(x) for x in el.get("rpy", "0 0 0").split()]
    return _rpy_matrix(*rpy), xyz


def _cylinder_mesh(radius: float, length: float, nseg: int = 24) -> Tuple[np.ndarray, np.ndarray]:
    ang = np.linspace(0, 2 * math.pi, nseg, endpoint=False)
    ring = np.stack([radius * np.cos(ang), radius * np.sin(ang)], axis=1)
    v = np.vstack([np.column_stack([ring, np.full(nseg, -length / 2)]), np.column_stack([ring, np.full(nseg, length / 2)]),
                   [[0, 0, -length / 2]], [[0, 0, length / 2]]])
    t = []
    for i in range(nseg):
        j = (i + 1) % nseg
        t += [[i, j, nseg + j], [i, nseg + j, nseg + i], [2 * nseg, j, i], [2 * nseg + 1, nseg + i, nseg + j]]
    return v, np.array(t, dtype=np.int32)


def load_urdf(path: str, world: Optional[WorldSpec] = None) -> Tuple[WorldSpec, RobotSpec]:
    return parse_urdf(open(path).read(), os.path.dirname(os.path.abspath(path)), world)


def parse_urdf(text: str, basedir: str = ".", world: Optional[WorldSpec] = None) -> Tuple[WorldSpec, RobotSpec]:
    """URDF -> RobotSpec the way RobotModel::LoadURDF builds its links (reference Cpp/Modeling/Robot.cpp:2566-3300):

    * links in depth-first order from the root; a root link called ``world`` (``<klampt world_frame>``) is dropped and the
      robot is fixed-base; any other root makes the robot floating: five virtual links base0..base4 (x, y, z prismatic; z, y
      revolute) precede the root link, which turns about x, under one Floating joint (:2864-2998) -- unless
      ``freeze_root_link``, which welds the six;
    * ``T0_Parent`` = the joint's ``<origin>``; axis = ``<axis>`` (default x); revolute -> Normal, continuous -> Spin (limits
      +-inf), prismatic -> Normal on a prismatic link, fixed -> Weld with limits 0 (:3086-3120); other joint types are errors;
    * ``<mimic>`` joints become affine drivers (:3121-3135, 3180-3230): q_child = multiplier * q_parent + offset;
    * collision geometry (visual with ``use_vis_geom``): meshes (OFF / OBJ / STL, ``package://`` relative to
      ``package_root``, ``scale``; ``flip_yz`` swaps the mesh's y and z as the reference's default does), boxes, cylinders
      and spheres tessellated; several ``<collision>`` elements of one link are merged; the element's ``<origin>`` is
      baked into the vertices;
    * ``<klampt><noselfcollision | selfcollision pairs= | group1= group2=>`` edit the default self-collision set."""
    import xml.etree.ElementTree as ET
    from . import synth
    world = world if world is not None else WorldSpec()
    root = ET.fromstring(text)
    kl = root.find("klampt")
    kattr = kl.attrib if kl is not None else {}
    flag = lambda k, d: (kattr.get(k, d).strip().lower() not in ("0", "false")) if isinstance(kattr.get(k, d), str) else bool(d)
    use_vis, flip_yz, freeze = flag("use_vis_geom", "0"), flag("flip_yz", "1"), flag("freeze_root_link", "0")
    world_frame, pkg_root = kattr.get("world_frame", "world"), kattr.get("package_root", ".")
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_of: Dict[str, list] = {}
    parent_joint = {}
    for j in joints:
        pa, ch = j.find("parent").get("link"), j.find("child").get("link")
        if ch not in links or pa not in links:
            raise ValueError("URDF joint %r references an unknown link" % j.get("name"))
        if ch in parent_joint:
            raise ValueError("URDF link %r has multiple parents" % ch)
        child_of.setdefault(pa, []).append(j)
        parent_joint[ch] = j
    roots = [n for n in links if n not in parent_joint]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)
    order, stack = [], [(roots[0], None)]
    while stack:                                              # depth-first, children in file order
        name, pj = stack.pop()
        order.append((name, pj))
        for j in reversed(child_of.get(name, [])):
            stack.append((j.find("child").get("link"), j))
    floating = roots[0] != world_frame
    names: List[str] = []
    parents: List[int] = []
    ltype: List[int] = []
    axes: List[List[float]] = []
    T0: List[np.ndarray] = []
    qmin: List[float] = []
    qmax: List[float] = []
    jt: List[int] = []
    jl: List[int] = []
    jb: List[int] = []
    geoms: List[int] = []
    index: Dict[str, int] = {}
    mimic = []

    def add(name, parent, lt, axis, T, lo, hi, g):
        names.append(name); parents.append(parent); ltype.append(lt); axes.append(list(axis)); T0.append(T)
        qmin.append(lo); qmax.append(hi); geoms.append(g)
        return len(names) - 1

    def link_geometry(lel) -> int:
        parts = []
        for c in lel.findall("visual" if use_vis else "collision"):
            g = c.find("geometry")
            if g is None or len(g) == 0:
                continue
            R, t = _origin(c.find("origin"))
            shape = g[0]
            if shape.tag == "mesh":
                fn = shape.get("filename")
                if fn.startswith("package://"):
                    fn = os.path.join(pkg_root, fn[len("package://"):])
                v, tr = load_mesh(fn if os.path.isabs(fn) else os.path.join(basedir, fn))
                if flip_yz:
                    v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], axis=1)      # y-up asset -> z-up (a rotation about x, no mirroring)
                v = v * np.array([float(x) for x in shape.get("scale", "1 1 1").split()], dtype=np.float64)
            elif shape.tag == "box":
                h = 0.5 * np.array([float(x) for x in shape.get("size").split()], dtype=np.float64)
                v, tr = synth.box_mesh(-h, h, div=2)
            elif shape.tag == "cylinder":
                v, tr = _cylinder_mesh(float(shape.get("radius")), float(shape.get("length")))
            elif shape.tag == "sphere":
                v, tr = synth.icosphere(2)
                v = v * float(shape.get("radius"))
            else:
                raise ValueError("URDF geometry <%s> is not supported" % shape.tag)
            parts.append((v @ R.T + t, tr))
        if not parts:
            return -1
        v, tr = synth.merge_meshes(parts)
        return world.add_geom(GeomSpec.mesh(v, tr))

    if floating:
        for i, (lt, ax) in enumerate(((PRISMATIC, (1, 0, 0)), (PRISMATIC, (0, 1, 0)), (PRISMATIC, (0, 0, 1)), (REVOLUTE, (0, 0, 1)), (REVOLUTE, (0, 1, 0)))):
            add("base%d" % i, i - 1, lt, ax, IDENTITY12.copy(), 0.0 if freeze else -np.inf, 0.0 if freeze else np.inf, -1)
        if freeze:
            for i in range(5):
                jt.append(JOINT_WELD); jl.append(i); jb.append(i - 1)
    for name, pj in order:
        lel = links[name]
        if pj is None:
            if not floating:
                continue                                      # the world link itself is not a robot link
            k = add(name, 4, REVOLUTE, (1, 0, 0), IDENTITY12.copy(), 0.0 if freeze else -np.inf, 0.0 if freeze else np.inf, link_geometry(lel))
            index[name] = k
            if freeze:
                jt.append(JOINT_WELD); jl.append(k); jb.append(4)
            else:
                jt.append(JOINT_FLOATING); jl.append(k); jb.append(-1)
            continue
        kind = pj.get("type")
        R, t = _origin(pj.find("origin"))
        T = np.concatenate([R.reshape(-1), t])
        ax_el = pj.find("axis")
        axis = [float(x) for x in ax_el.get("xyz").split()] if ax_el is not None else [1.0, 0.0, 0.0]
        n = math.sqrt(sum(a * a for a in axis))
        if kind != "fixed" and n < 0.1:
            raise ValueError("URDF joint %r has a degenerate axis" % pj.get("name"))
        axis = [a / n for a in axis] if n > 0 else [1.0, 0.0, 0.0]
        lim = pj.find("limit")
        lo, hi = (-np.inf, np.inf) if lim is None else (float(lim.get("lower", "0")), float(lim.get("upper", "0")))
        if kind == "revolute":
            jtype, lt = JOINT_NORMAL, REVOLUTE
        elif kind == "continuous":
            jtype, lt, lo, hi = JOINT_SPIN, REVOLUTE, -np.inf, np.inf
        elif kind == "prismatic":
            jtype, lt = JOINT_NORMAL, PRISMATIC
        elif kind == "fixed":
            jtype, lt, lo, hi = JOINT_WELD, REVOLUTE, 0.0, 0.0
        else:
            raise ValueError("URDF joint type %r is not supported" % kind)
        parent_name = pj.find("parent").get("link")
        par = index.get(parent_name, -1)
        k = add(name, par, lt, axis, T, lo, hi, link_geometry(lel))
        index[name] = k
        jt.append(jtype); jl.append(k); jb.append(par)
        m = pj.find("mimic")
        if m is not None:
            mimic.append((k, m.get("joint"), float(m.get("multiplier", "1")), float(m.get("offset", "0"))))
    L = len(names)
    if L == 0:
        raise ValueError("URDF has no links besides the world frame")
    jname_to_link = {j.get("name"): index[j.find("child").get("link")] for j in joints if j.find("child").get("link") in index}
    drivers = []
    for k, jn, mult, off in mimic:
        if jn not in jname_to_link:
            continue
        src = jname_to_link[jn]
        # q_k = mult * q_src + off: an affine driver over (src, k) with scales (1, mult) and offsets (0, off)
        drivers.append(DriverSpec([src, k], [1.0, mult], [0.0, off], float(qmin[src]), float(qmax[src])))

    def resolve(tok: str) -> int:
        try:
            return int(tok)
        except ValueError:
            return names.index(tok)

    edits = []
    if kl is not None:
        if kl.findall("selfcollision"):
            edits += [(i, j, False) for i in range(L) for j in range(i + 1, L)]
        for tag, en in (("selfcollision", True), ("noselfcollision", False)):
            for el in kl.findall(tag):
                pairs = []
                if el.get("pairs"):
                    toks = el.get("pairs").split()
                    if len(toks) % 2:
                        raise ValueError("<%s pairs> holds link PAIRS" % tag)
                    pairs = list(zip(toks[0::2], toks[1::2]))
                elif el.get("group1") and el.get("group2"):
                    pairs = [(a, b) for a in el.get("group1").split() for b in el.get("group2").split()]
                else:
                    raise ValueError("<%s> needs pairs or group1 + group2" % tag)
                for a, b in pairs:
                    i, j = resolve(a), resolve(b)
                    if i != j:
                        edits.append((min(i, j), max(i, j), en))
    spec = RobotSpec(parents=np.array(parents, dtype=np.int32), linktype=np.array(ltype, dtype=np.uint8), axis=np.array(axes, dtype=np.float64),
                     T0=np.array(T0, dtype=np.float64), qmin=np.array(qmin, dtype=np.float64), qmax=np.array(qmax, dtype=np.float64),
                     link_geom=geoms, joint_type=np.array(jt, dtype=np.uint8), joint_link=np.array(jl, dtype=np.int32),
                     joint_base=np.array(jb, dtype=np.int32), drivers=drivers, self_collision_edits=edits, names=names)
    world.robot = spec
    return world, spec


# ============================================================================================== world XML
def _vec(txt: Optional[str], n: int, default: float = 0.0) -> np.ndarray:
    if txt is None:
        return np.full(n, default, dtype=np.float64)
    v = np.array([float(x) for x in txt.split()], dtype=np.float64)
    return np.full(n, v[0]) if len(v) == 1 else v[:n]


def _xml_rotation(el) -> np.ndarray:
    """ReadTransform of the world loader (Cpp/IO/XmlWorld.cpp:97-160): rotateRPY sets the rotation, rotateMoment replaces it, then
    rotateX, rotateY, rotateZ are pre-multiplied IN THAT FIXED ORDER -- whatever order the attributes are written in"""
    from . import so3
    R = np.eye(3)
    if el.get("rotateRPY") is not None:
        r, p, y = _vec(el.get("rotateRPY"), 3)
        R = _rpy_matrix(r, p, y)
    if el.get("rotateMoment") is not None:
        R = so3.exp(_vec(el.get("rotateMoment"), 3))
    for key, ax in (("rotateX", [1, 0, 0]), ("rotateY", [0, 1, 0]), ("rotateZ", [0, 0, 1])):
        if el.get(key) is not None:
            R = so3.matrix(so3.from_axis_angle((ax, float(el.get(key))))) @ R
    return R


def _xml_translation(el, extra=()) -> np.ndarray:
    """translation, else position (ReadTransform, XmlWorld.cpp:103-107); `extra`: further attribute names accepted after those"""
    for key in ("translation", "position") + tuple(extra):
        if el.get(key) is not None:
            return _vec(el.get(key), 3)
    return np.zeros(3)


def load_world_xml(path: str) -> WorldSpec:
    return parse_world_xml(open(path).read(), os.path.dirname(os.path.abspath(path)))


def parse_world_xml(text: str, basedir: str = ".") -> WorldSpec:
    """The entities of a world file that the feasibility path reads (format: Cpp/docs/Manual-FileTypes.md:47-162; loader:
    Cpp/IO/XmlWorld.cpp): ``<robot file=.rob|.urdf>`` (the first one is the engine's active robot, any further robot is
    rejected -- one active robot per engine), ``<rigidObject>`` with ``<geometry file|mesh scale translate margin>`` and
    ``position`` / ``rotate*``, ``<terrain file scale margin translation|position rotate*>``.  Display, physics and
    simulation elements are parsed over.  A terrain's pose is baked into its geometry (terrains carry no transform)."""
    import xml.etree.ElementTree as ET
    root = ET.fromstring(text)
    if root.tag != "world":
        raise ValueError("not a world file: top-level element is <%s>" % root.tag)
    w = WorldSpec()
    resolve = lambda fn: fn if os.path.isabs(fn) else os.path.join(basedir, fn)
    for el in root:
        if el.tag == "terrain":
            g = load_geometry(resolve(el.get("file")), _vec(el.get("scale"), 3, 1.0), (0, 0, 0), float(el.get("margin", "0")))
            R, t = _xml_rotation(el), _xml_translation(el)
            if g.kind == "mesh":
                g.verts = g.verts @ R.T + t
            else:
                g.points = g.points @ R.T + t
            w.terrains.append(w.add_geom(g))
        elif el.tag == "rigidObject":
            ge = el.find("geometry")
            if ge is None:
                raise ValueError("<rigidObject> without <geometry> (rigid object .obj description files are not read)")
            fn = ge.get("file", ge.get("mesh"))
            # <geometry> carries its own ReadTransform (XmlWorld.cpp:292-294): p -> R (s * p) + t, applied to the geometry itself
            g = load_geometry(resolve(fn), _vec(ge.get("scale"), 3, 1.0), (0.0, 0.0, 0.0), float(ge.get("margin", "0")))
            Rg, tg = _xml_rotation(ge), _xml_translation(ge, extra=("translate",))
            if g.kind == "mesh":
                g.verts = g.verts @ Rg.T + tg
            else:
                g.points = g.points @ Rg.T + tg
            T = np.concatenate([_xml_rotation(el).reshape(-1), _xml_translation(el)])
            w.objects.append((w.add_geom(g), T))
        elif el.tag == "robot":
            if w.robot is not None:
                raise ValueError("more than one <robot>: the engine has one active robot (SingleRobotCSpace)")
            fn = resolve(el.get("file"))
            if fn.lower().endswith(".urdf"):
                load_urdf(fn, w)
            else:
                load_rob(fn, w)
            # XmlRobot::GetRobot (XmlWorld.cpp:251-257): the element's transform is pre-multiplied into T0_Parent of every root link
            Rr, tr = _xml_rotation(el), _xml_translation(el)
            if not (np.array_equal(Rr, np.eye(3)) and not tr.any()):
                r = w.robot
                for i in range(r.L):
                    if r.parents[i] < 0:
                        R0, t0 = r.T0[i, :9].reshape(3, 3).copy(), r.T0[i, 9:12].copy()
                        r.T0[i, :9] = (Rr @ R0).reshape(-1)
                        r.T0[i, 9:12] = Rr @ t0 + tr
    return w
