"""Python face of the C ABI: builds one engine handle from a WorldSpec and exposes the batched hot path.

Host entry points take / return numpy arrays (H2D and D2H copies happen inside the library).  The
``*_device`` entry points take raw device pointers (ints, or anything with ``data_ptr()`` such as a torch
tensor) and enqueue on the engine's stream: PyTorch is only plumbing for device memory and streams here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _capi
from ._capi import KbStats, check
from .worldspec import WorldSpec


def _ptr(x):
    """device / host pointer of a numpy array, torch tensor, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError("cannot take a pointer of %r" % type(x))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


class Engine:
    """One kb_engine handle = static world + one active robot, resident on one GPU."""

    def __init__(self, spec: WorldSpec, device=0, options: Optional[dict] = None):
        """device: one CUDA device index, or a list of them (kb_finalize_multi: host-buffer batches are sharded over the devices).
        options: kb_set_option values that must be in place before kb_finalize, e.g. {"cloud_builder": 1} (point-cloud
        hierarchies built on the GPU) or {"grid_res": 0}"""
        self.lib = _capi.load()
        self.spec = spec
        self.h = C.c_void_p()
        check(self.lib.kb_engine_create(C.byref(self.h)))
        try:
            self._describe(spec)
            options = dict(options or {})
            # experiment knob: KLAMPT_B200_OPTIONS="name=value,name=value" is applied to every engine of the process
            for kv in filter(None, os.environ.get("KLAMPT_B200_OPTIONS", "").split(",")):
                k, v = kv.split("=")
                options[k.strip()] = int(v)
            for k, v in options.items():
                check(self.lib.kb_set_option(self.h, k.encode(), int(v)))
            if isinstance(device, (list, tuple)):
                devs = np.ascontiguousarray(device, dtype=np.int32)
                check(self.lib.kb_finalize_multi(self.h, devs.ctypes.data_as(_capi.c_int32_p), len(devs)))
                device = int(devs[0])
            else:
                check(self.lib.kb_finalize(self.h, int(device)))
        except Exception:
            self.close()
            raise
        self.device = int(device)
        self.L = spec.robot.L

    # ------------------------------------------------------------------ construction
    def _describe(self, spec: WorldSpec):
        lib, h = self.lib, self.h
        dp, ip, up = _capi.c_double_p, _capi.c_int32_p, _capi.c_uint8_p
        for g in spec.geoms:
            if g.kind == "mesh":
                v = _f64(g.verts)
                t = np.ascontiguousarray(g.tris, dtype=np.int32)
                check(lib.kb_add_trimesh(h, v.ctypes.data_as(dp), len(v), t.ctypes.data_as(ip), len(t), g.margin))
            elif g.kind == "cloud":
                p = _f64(g.points)
                r = None if g.radius is None else _f64(g.radius)
                check(lib.kb_add_pointcloud(h, p.ctypes.data_as(dp), len(p), None if r is None else r.ctypes.data_as(dp), g.margin))
            elif g.kind == "dyncloud":
                check(lib.kb_add_dynamic_pointcloud(h, int(g.params[0]), float(g.params[1]), g.margin))
            elif g.kind in ("triangle", "box", "segment"):
                p = _f64(g.params)
                check(lib.kb_add_primitive(h, {"triangle": 2, "box": 3, "segment": 5}[g.kind], p.ctypes.data_as(dp), g.margin))
            elif g.kind in ("sphere", "point"):
                p = _f64(g.params)
                check(lib.kb_add_primitive(h, 1 if g.kind == "sphere" else 0, p.ctypes.data_as(dp), g.margin))
            elif g.kind == "empty":
                check(lib.kb_add_trimesh(h, None, 0, None, 0, 0.0))
            else:
                raise ValueError("unsupported geometry kind %r" % g.kind)
        for gi in spec.terrains:
            check(lib.kb_add_terrain(h, int(gi)))
        for gi, T in spec.objects:
            T = _f64(T, (12,))
            check(lib.kb_add_rigid_object(h, int(gi), T.ctypes.data_as(dp)))
        r = spec.robot
        if r is None:
            raise ValueError("WorldSpec has no robot")
        par = np.ascontiguousarray(r.parents, dtype=np.int32)
        lt = np.ascontiguousarray(r.linktype, dtype=np.uint8)
        ax, T0, qmin, qmax = _f64(r.axis), _f64(r.T0), _f64(r.qmin), _f64(r.qmax)
        check(lib.kb_robot_create(h, r.L, par.ctypes.data_as(ip), lt.ctypes.data_as(up), ax.ctypes.data_as(dp), T0.ctypes.data_as(dp),
                                  qmin.ctypes.data_as(dp), qmax.ctypes.data_as(dp)))
        for j, gi in enumerate(r.link_geom):
            check(lib.kb_robot_set_link_geometry(h, j, int(gi)))
        if r.joint_type is not None:
            jt = np.ascontiguousarray(r.joint_type, dtype=np.uint8)
            jl = np.ascontiguousarray(r.joint_link, dtype=np.int32)
            jb = None if getattr(r, "joint_base", None) is None else np.ascontiguousarray(r.joint_base, dtype=np.int32)
            check(lib.kb_robot_set_joints(h, len(jt), jt.ctypes.data_as(up), jl.ctypes.data_as(ip), None if jb is None else jb.ctypes.data_as(ip)))
        for d in r.drivers:
            li = np.ascontiguousarray(d.links, dtype=np.int32)
            sc, of = _f64(d.scale), _f64(d.offset)
            check(lib.kb_robot_add_driver(h, len(li), li.ctypes.data_as(ip), sc.ctypes.data_as(dp), of.ctypes.data_as(dp), d.qmin, d.qmax))
        for (i, j, en) in r.self_collision_edits:
            check(lib.kb_robot_set_self_collision(h, int(i), int(j), int(bool(en))))
        if spec.pair_mask is not None:
            m = np.ascontiguousarray(spec.pair_mask, dtype=np.uint8)
            check(lib.kb_set_pair_mask(h, m.ctypes.data_as(up), m.shape[0]))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.kb_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def set_stream(self, stream_ptr: Optional[int]):
        """stream_ptr: a cudaStream_t as int (e.g. torch.cuda.current_stream().cuda_stream).  0 means CUDA's legacy
        default stream (passed to the library as cudaStreamLegacy); None restores the engine's own stream."""
        if stream_ptr is None:
            check(self.lib.kb_set_stream(self.h, None))
        else:
            check(self.lib.kb_set_stream(self.h, C.c_void_p(stream_ptr if stream_ptr else 1)))

    def synchronize(self):
        check(self.lib.kb_synchronize(self.h))

    def set_option(self, name: str, value: int):
        check(self.lib.kb_set_option(self.h, name.encode(), int(value)))

    def num_devices(self) -> int:
        return int(self.lib.kb_num_devices(self.h))

    def num_ids(self) -> int:
        return int(self.lib.kb_num_ids(self.h))

    def pair_mask(self) -> np.ndarray:
        n = self.num_ids()
        m = np.zeros((n, n), dtype=np.uint8)
        check(self.lib.kb_get_pair_mask(self.h, m.ctypes.data_as(_capi.c_uint8_p)))
        return m

    def stats(self) -> dict:
        s = KbStats()
        check(self.lib.kb_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in KbStats._fields_}

    def reset_stats(self):
        check(self.lib.kb_reset_stats(self.h))

    def layout(self) -> dict:
        a = (C.c_int64 * 6)()
        check(self.lib.kb_get_layout(self.h, a))
        return {"nodes": a[0], "elements": a[1], "static_bytes": a[2], "items_per_config": a[3], "env_groups": a[4], "max_depth_sum": a[5]}

    # ------------------------------------------------------------------ hot path, host buffers
    def _Q(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.ndim == 1:
            Q = Q.reshape(1, -1)
        if Q.ndim != 2 or Q.shape[1] != self.L:
            raise ValueError("configurations must be (N, %d), got %r" % (self.L, Q.shape))
        return Q

    def fk_batch(self, Q) -> np.ndarray:
        Q = self._Q(Q)
        T = np.empty((Q.shape[0], self.L, 12), dtype=np.float64)
        check(self.lib.kb_fk_batch(self.h, _ptr(Q), Q.shape[0], _ptr(T)))
        return T

    def feasible_batch(self, Q, return_pairs: bool = False, out: Optional[np.ndarray] = None):
        if isinstance(Q, np.ndarray) and Q.dtype == np.float32:      # fp32 rows travel as they are and are widened on the device
            Qf = np.ascontiguousarray(Q).reshape(-1, self.L)
            N = Qf.shape[0]
            if out is None:
                out = np.empty(N, dtype=np.uint8)
            pairs = np.empty((N, 2), dtype=np.int32) if return_pairs else None
            check(self.lib.kb_feasible_batch_f32(self.h, _ptr(Qf), N, _ptr(out), _ptr(pairs)))
            return (out, pairs) if return_pairs else out
        Q = self._Q(Q)
        N = Q.shape[0]
        if out is None:
            out = np.empty(N, dtype=np.uint8)
        pairs = np.empty((N, 2), dtype=np.int32) if return_pairs else None
        check(self.lib.kb_feasible_batch(self.h, _ptr(Q), N, _ptr(out), _ptr(pairs)))
        return (out, pairs) if return_pairs else out

    def feasible_batch_bits(self, Q) -> np.ndarray:
        """feasibility as a packed bitmask, (N + 7) // 8 bytes, little bit order: np.unpackbits(bits, bitorder="little")[:N]"""
        Q = self._Q(Q)
        N = Q.shape[0]
        bits = np.zeros((N + 7) // 8, dtype=np.uint8)
        check(self.lib.kb_feasible_batch_bits(self.h, _ptr(Q), N, _ptr(bits)))
        return bits

    def edges_visible_batch_bits(self, A, B, eps: float = 0.01, weights=None) -> np.ndarray:
        A, B = self._Q(A), self._Q(B)
        N = A.shape[0]
        bits = np.zeros((N + 7) // 8, dtype=np.uint8)
        w = None if weights is None else _f64(weights)
        check(self.lib.kb_edges_visible_batch_bits(self.h, _ptr(A), _ptr(B), N, float(eps), _ptr(w), _ptr(bits), None))
        return bits

    def edges_visible_batch(self, A, B, eps: float = 0.01, weights=None, return_nchecks: bool = True):
        A, B = self._Q(A), self._Q(B)
        if A.shape != B.shape:
            raise ValueError("A and B must have the same shape")
        N = A.shape[0]
        out = np.empty(N, dtype=np.uint8)
        nchecks = np.empty(N, dtype=np.int32) if return_nchecks else None
        w = None if weights is None else _f64(weights)
        check(self.lib.kb_edges_visible_batch(self.h, _ptr(A), _ptr(B), N, float(eps), _ptr(w), _ptr(out), _ptr(nchecks)))
        return (out, nchecks) if return_nchecks else out

    def distance_batch(self, Q, upper_bound: float = np.inf, include_self: bool = False, return_pairs: bool = False):
        Q = self._Q(Q)
        N = Q.shape[0]
        d = np.empty(N, dtype=np.float64)
        pairs = np.empty((N, 2), dtype=np.int32) if return_pairs else None
        check(self.lib.kb_distance_batch(self.h, _ptr(Q), N, float(upper_bound), int(include_self), _ptr(d), _ptr(pairs)))
        return (d, pairs) if return_pairs else d

    def distance_batch_ex(self, Q, upper_bound: float = np.inf, include_self: bool = False, abs_err: float = 0.0, rel_err: float = 0.0):
        """AnyCollisionQuery::Distance(absErr, relErr, bound) with the whole DistanceQueryResult: (d (N,), pairs (N, 2) world ids,
        closest points (N, 2, 3) on the margin-inflated surfaces -- [:, 0] on pairs[:, 0]'s geometry --, element indices (N, 2))"""
        Q = self._Q(Q)
        N = Q.shape[0]
        d = np.empty(N, dtype=np.float64)
        pairs = np.empty((N, 2), dtype=np.int32)
        cp = np.empty((N, 2, 3), dtype=np.float64)
        elem = np.empty((N, 2), dtype=np.int32)
        check(self.lib.kb_distance_batch_ex(self.h, _ptr(Q), N, float(abs_err), float(rel_err), float(upper_bound), int(include_self),
                                            _ptr(d), _ptr(pairs), _ptr(cp), _ptr(elem)))
        return d, pairs, cp, elem

    def geom_distance_batch_ex(self, ga: int, Ta, gb: int, Tb, upper_bound: float = np.inf, abs_err: float = 0.0, rel_err: float = 0.0):
        """Geometry3D.distance_ext for N transform pairs: (d, closest points (N, 2, 3): [:, 0] on ga, [:, 1] on gb, element indices (N, 2))"""
        Ta, Tb = _f64(Ta).reshape(-1, 12), _f64(Tb).reshape(-1, 12)
        N = Ta.shape[0]
        d = np.empty(N, dtype=np.float64)
        cp = np.empty((N, 2, 3), dtype=np.float64)
        elem = np.empty((N, 2), dtype=np.int32)
        check(self.lib.kb_geom_distance_batch_ex(self.h, int(ga), _ptr(Ta), int(gb), _ptr(Tb), N, float(abs_err), float(rel_err), float(upper_bound),
                                                 _ptr(d), _ptr(cp), _ptr(elem)))
        return d, cp, elem

    def _ignore_mask(self, ignore_ids) -> np.ndarray:
        ig = np.asarray(ignore_ids)
        nids = self.num_ids()
        if ig.dtype != np.uint8 or ig.shape != (nids,):
            m = np.zeros(nids, dtype=np.uint8)
            m[np.asarray(list(ignore_ids), dtype=np.int64)] = 1
            ig = m
        return np.ascontiguousarray(ig)

    def raycast_batch(self, q, rays, ignore_ids=None):
        """WorldModel::RayCast / RayCastIgnore (World.cpp:465-588) for N rays (rows of source xyz, direction xyz) with the robot at q
        (None: the robot is left out): (world id or -1, distance along the normalised direction or inf, element index or -1).
        ignore_ids: world ids the rays pass through (an iterable of ids, or one byte per id)."""
        f32 = isinstance(rays, np.ndarray) and rays.dtype == np.float32       # fp32 rays travel as they are and are widened on the device
        rays = np.ascontiguousarray(rays).reshape(-1, 6) if f32 else _f64(rays).reshape(-1, 6)
        N = rays.shape[0]
        ids, dist, elem = np.empty(N, dtype=np.int32), np.empty(N, dtype=np.float64), np.empty(N, dtype=np.int32)
        qa = None if q is None else _f64(q).reshape(self.L)
        ig = None if ignore_ids is None else self._ignore_mask(ignore_ids)
        check((self.lib.kb_raycast_batch_f32 if f32 else self.lib.kb_raycast_batch)(self.h, None if qa is None else _ptr(qa), _ptr(rays), N, None if ig is None else _ptr(ig),
                                        _ptr(ids), _ptr(dist), _ptr(elem)))
        return ids, dist, elem

    def camera_depth(self, q, pose, fx, fy, cx, cy, zmin, zmax, xres, yres, ignore_ids=None, want_ids=True):
        """CameraSensor's ray-cast rendering in one call (rays built on the device): (depth (yres, xres) float32, world ids (yres, xres)
        int32 or None).  pose: the camera's world pose (12 doubles; x right, y down, z forward)."""
        from ._capi import KbCamera
        cam = KbCamera()
        cam.pose[:] = list(_f64(pose).reshape(12))
        cam.fx, cam.fy, cam.cx, cam.cy, cam.zmin, cam.zmax, cam.xres, cam.yres = fx, fy, cx, cy, zmin, zmax, int(xres), int(yres)
        depth = np.empty((int(yres), int(xres)), dtype=np.float32)
        ids = np.empty((int(yres), int(xres)), dtype=np.int32) if want_ids else None
        qa = None if q is None else _f64(q).reshape(self.L)
        ig = None if ignore_ids is None else self._ignore_mask(ignore_ids)
        check(self.lib.kb_camera_depth(self.h, None if qa is None else _ptr(qa), C.byref(cam), None if ig is None else _ptr(ig), _ptr(depth),
                                       None if ids is None else _ptr(ids)))
        return depth, ids

    def geom_raycast_batch(self, geom: int, T, rays):
        """Geometry3D.rayCast_ext of one registered geometry at transform T (12 doubles or None) for N rays: (element or -1, distance or inf)"""
        rays = _f64(rays).reshape(-1, 6)
        N = rays.shape[0]
        elem, dist = np.empty(N, dtype=np.int32), np.empty(N, dtype=np.float64)
        Ta = None if T is None else _f64(T).reshape(12)
        check(self.lib.kb_geom_raycast_batch(self.h, int(geom), None if Ta is None else _ptr(Ta), _ptr(rays), N, _ptr(elem), _ptr(dist)))
        return elem, dist

    def colliding_pairs_batch(self, Q, max_pairs: int = 8):
        """every colliding (idA, idB) world-id pair per configuration: (pairs (N, max_pairs, 2) padded with -1, count (N,));
        count = -1 where the joint / driver limits already fail"""
        Q = self._Q(Q)
        N = Q.shape[0]
        pairs = np.empty((N, max_pairs, 2), dtype=np.int32)
        count = np.empty(N, dtype=np.int32)
        check(self.lib.kb_colliding_pairs_batch(self.h, _ptr(Q), N, int(max_pairs), _ptr(pairs), _ptr(count)))
        return pairs, count

    def geom_collides_batch(self, ga: int, Ta, gb: int, Tb, tol: float = 0.0) -> np.ndarray:
        Ta, Tb = _f64(Ta).reshape(-1, 12), _f64(Tb).reshape(-1, 12)
        N = Ta.shape[0]
        out = np.empty(N, dtype=np.uint8)
        check(self.lib.kb_geom_collides_batch(self.h, int(ga), _ptr(Ta), int(gb), _ptr(Tb), N, float(tol), _ptr(out)))
        return out

    def geom_distance_batch(self, ga: int, Ta, gb: int, Tb, upper_bound: float = np.inf) -> np.ndarray:
        Ta, Tb = _f64(Ta).reshape(-1, 12), _f64(Tb).reshape(-1, 12)
        N = Ta.shape[0]
        out = np.empty(N, dtype=np.float64)
        check(self.lib.kb_geom_distance_batch(self.h, int(ga), _ptr(Ta), int(gb), _ptr(Tb), N, float(upper_bound), _ptr(out)))
        return out

    def update_pointcloud(self, geom: int, points) -> None:
        """Geometry3D.setPointCloud on a dynamic cloud (GeomSpec.dynamic_cloud): uploads the points (local frame) and rebuilds
        the cloud's hierarchy on the GPU"""
        p = _f64(points).reshape(-1, 3)
        check(self.lib.kb_update_pointcloud(self.h, int(geom), _ptr(p) if len(p) else None, len(p)))

    # ------------------------------------------------------------------ hot path, device buffers
    def feasible_batch_device(self, dQ, N: int, d_out, d_first_pair=None):
        check(self.lib.kb_feasible_batch_device(self.h, _ptr(dQ), int(N), _ptr(d_out), _ptr(d_first_pair)))

    def feasible_batch_bits_device(self, dQ, N: int, d_bits):
        check(self.lib.kb_feasible_batch_bits_device(self.h, _ptr(dQ), int(N), _ptr(d_bits)))

    def edges_visible_batch_device(self, dA, dB, N: int, eps: float, d_out, d_nchecks=None, weights=None):
        w = None if weights is None else _f64(weights)
        check(self.lib.kb_edges_visible_batch_device(self.h, _ptr(dA), _ptr(dB), int(N), float(eps), _ptr(w), _ptr(d_out), _ptr(d_nchecks)))

    def distance_batch_device(self, dQ, N: int, upper_bound: float, include_self: bool, d_out_d, d_out_pair=None):
        check(self.lib.kb_distance_batch_device(self.h, _ptr(dQ), int(N), float(upper_bound), int(include_self), _ptr(d_out_d), _ptr(d_out_pair)))
