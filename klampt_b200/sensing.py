"""Kinematic simulation of the two ray-cast sensors on the batched ray caster (``Engine.raycast_batch``).

The reference simulates a ``CameraSensor`` without OpenGL, and every ``LaserRangeSensor``, by calling ``WorldModel::RayCast`` /
``RayCastIgnore`` once per pixel / measurement (Cpp/Sensing/VisualSensors.cpp:413-475 and :57-142).  Here the rays of one
reading are built the same way and cast in one launch.  Restated from those two functions:

* camera: viewport from ``CameraSensor::GetViewport`` (:865-892; fx = xres / 2 / tan(xfov / 2) unless given, cx = xres / 2, ...),
  camera frame x right, y down, z forward (Klamp't's sensor convention); pixel (i, j) looks along
  ``fwd + (i - cx) right / fx + (cy - j) up / fy``, the ray starts ``zmin`` along that unnormalised vector, depth =
  ``fwd . (pt - eye)``, readings below ``zmin`` or beyond ``zmax`` become ``zmax``, misses become ``zmax``;
* laser: measurement i looks along ``(sin x, cos x sin y, cos x cos y)`` in the sensor frame with the sweep angles x, y of
  ``EvalPattern``; the ray starts ``depthMinimum`` along it, the reading is the distance from there plus ``depthMinimum``; the
  link the sensor is mounted on is ignored; readings at or below ``depthMinimum`` or at or beyond ``depthMaximum`` become ``depthMaximum``.

Noise and discretisation (``Discretize`` with the variance terms) are left to the caller: they are per-reading host arithmetic.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

SWEEP_SINUSOID, SWEEP_TRIANGULAR, SWEEP_SAWTOOTH = 0, 1, 2


def _T(T12) -> Tuple[np.ndarray, np.ndarray]:
    T12 = np.asarray(T12, dtype=np.float64).reshape(12)
    return T12[:9].reshape(3, 3), T12[9:]


class CameraSensor:
    """the ray-cast path of CameraSensor::SimulateKinematic.  Tsensor: pose of the camera in its link's frame (12 doubles, row-major
    R then t); link: index of the link it rides on, -1 = the world frame."""

    def __init__(self, xres=640, yres=480, xfov=math.radians(56.0), yfov=math.radians(43.0), zmin=0.4, zmax=4.0, link=-1, Tsensor=None,
                 fx=-1.0, fy=-1.0, cx=-1.0, cy=-1.0):
        self.xres, self.yres, self.xfov, self.yfov, self.zmin, self.zmax, self.link = int(xres), int(yres), xfov, yfov, zmin, zmax, int(link)
        self.Tsensor = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64) if Tsensor is None else np.asarray(Tsensor, dtype=np.float64)
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy

    def viewport(self):
        fx = self.xres * 0.5 / math.tan(self.xfov * 0.5) if self.fx < 0 else self.fx
        fy = self.yres * 0.5 / math.tan(self.yfov * 0.5) if self.fy < 0 else self.fy
        cx = self.xres * 0.5 if self.cx < 0 else self.cx
        cy = self.yres * 0.5 if self.cy < 0 else self.cy
        return fx, fy, cx, cy

    def pose(self, link_transforms=None) -> Tuple[np.ndarray, np.ndarray]:
        R, t = _T(self.Tsensor)
        if self.link >= 0:
            Rl, tl = _T(np.asarray(link_transforms, dtype=np.float64).reshape(-1, 12)[self.link])
            R, t = Rl @ R, Rl @ t + tl
        return R, t

    def rays(self, link_transforms=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(rays (yres * xres, 6), eye, forward): row k = j * xres + i, as the reference fills its image"""
        R, eye = self.pose(link_transforms)
        right, up, fwd = R[:, 0], -R[:, 1], R[:, 2]
        fx, fy, cx, cy = self.viewport()
        u = (np.arange(self.xres, dtype=np.float64) - cx)[None, :, None]
        v = (cy - np.arange(self.yres, dtype=np.float64))[:, None, None]
        d = fwd[None, None, :] + u * (right * (1.0 / fx))[None, None, :] + v * (up * (1.0 / fy))[None, None, :]
        src = eye[None, None, :] + d * self.zmin
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        return np.concatenate([src, d], axis=-1).reshape(-1, 6), eye, fwd

    def simulate(self, engine, q, ignore_ids=None):
        """(depth image (yres, xres) float32, id image (yres, xres) int32: world id seen by each pixel, -1 = background); the rays are
        built on the device (kb_camera_depth): nothing but the pose goes up, only the two images come back"""
        T = engine.fk_batch(np.asarray(q, dtype=np.float64)[None, :])[0] if self.link >= 0 else None
        R, eye = self.pose(T)
        fx, fy, cx, cy = self.viewport()
        return engine.camera_depth(q, np.concatenate([R.reshape(-1), eye]), fx, fy, cx, cy, self.zmin, self.zmax, self.xres, self.yres, ignore_ids)

    def simulate_from_rays(self, engine, q, ignore_ids=None):
        """the same reading with the rays built on the host and cast through kb_raycast_batch (cross-check of the device-side builder)"""
        T = engine.fk_batch(np.asarray(q, dtype=np.float64)[None, :])[0] if self.link >= 0 else None
        rays, eye, fwd = self.rays(T)
        ids, dist, _ = engine.raycast_batch(q, rays, ignore_ids)
        hit = ids >= 0
        pt = rays[:, :3] + np.where(hit, dist, 0.0)[:, None] * rays[:, 3:]
        d = (pt - eye) @ fwd
        d = np.where(d < self.zmin, self.zmax, np.minimum(d, self.zmax))
        depth = np.where(hit, d, self.zmax).astype(np.float32)
        return depth.reshape(self.yres, self.xres), ids.reshape(self.yres, self.xres)


def eval_pattern(kind: int, x, correction: float = 1.0):
    x = np.asarray(x, dtype=np.float64)
    if kind == SWEEP_SINUSOID:
        return np.sin(x * 2.0 * math.pi)
    if kind == SWEEP_TRIANGULAR:
        return 2.0 * (1.0 + np.abs(np.mod(x, 2.0) - 1.0)) - 1.0          # as written in the reference (VisualSensors.cpp:52)
    return 2.0 * (np.mod(x / correction, 1.0) * correction) - 1.0


class LaserRangeSensor:
    """LaserRangeSensor::SimulateKinematic for one reading (measurementCount rays)"""

    def __init__(self, measurementCount=180, depthMinimum=0.1, depthMaximum=float("inf"), xSweepMagnitude=math.radians(90.0), xSweepPeriod=0.0,
                 xSweepPhase=0.0, xSweepType=SWEEP_SAWTOOTH, ySweepMagnitude=0.0, ySweepPeriod=0.0, ySweepPhase=0.0, ySweepType=SWEEP_SINUSOID,
                 link=-1, Tsensor=None):
        self.measurementCount, self.depthMinimum, self.depthMaximum = int(measurementCount), depthMinimum, depthMaximum
        self.xSweepMagnitude, self.xSweepPeriod, self.xSweepPhase, self.xSweepType = xSweepMagnitude, xSweepPeriod, xSweepPhase, xSweepType
        self.ySweepMagnitude, self.ySweepPeriod, self.ySweepPhase, self.ySweepType = ySweepMagnitude, ySweepPeriod, ySweepPhase, ySweepType
        self.link = int(link)
        self.Tsensor = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64) if Tsensor is None else np.asarray(Tsensor, dtype=np.float64)
        self.last_t, self.last_dt = 0.0, 0.0

    def advance(self, dt: float):
        self.last_dt = dt
        self.last_t += dt

    def angles(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.measurementCount
        xscale = 1.0
        if (self.xSweepType == SWEEP_SAWTOOTH or self.ySweepType == SWEEP_SAWTOOTH) and self.last_dt > 0 and n > 1:
            xscale = 1.0 + 1.0 / (n - 1)                 # the reference scales x for either sweep (its y branch assigns xscale too)
        ux0 = 0.0 if self.xSweepPeriod == 0 else (self.last_t - self.last_dt + self.xSweepPhase) / self.xSweepPeriod
        ux1 = 1.0 if self.xSweepPeriod == 0 else (self.last_t + self.xSweepPhase) / self.xSweepPeriod
        uy0 = 0.0 if self.ySweepPeriod == 0 else (self.last_t - self.last_dt + self.ySweepPhase) / self.ySweepPeriod
        uy1 = 1.0 if self.ySweepPeriod == 0 else (self.last_t + self.ySweepPhase) / self.ySweepPeriod
        if self.xSweepPeriod != 0 and n > 1:
            ux0 += (ux1 - ux0) / (n - 1)
        if self.ySweepPeriod != 0 and n > 1:
            uy0 += (uy1 - uy0) / (n - 1)
        step = 1.0 / (n - 1) if n > 1 else 0.0
        i = np.arange(n, dtype=np.float64)
        ux, uy = ux0 + i * step * (ux1 - ux0), uy0 + i * step * (uy1 - uy0)
        if n > 0:
            ux[-1], uy[-1] = ux1, uy1
        return self.xSweepMagnitude * eval_pattern(self.xSweepType, ux, xscale), self.ySweepMagnitude * eval_pattern(self.ySweepType, uy, 1.0)

    def rays(self, link_transforms=None) -> np.ndarray:
        R, t = _T(self.Tsensor)
        if self.link >= 0:
            Rl, tl = _T(np.asarray(link_transforms, dtype=np.float64).reshape(-1, 12)[self.link])
            R, t = Rl @ R, Rl @ t + tl
        xt, yt = self.angles()
        local = np.stack([np.sin(xt), np.cos(xt) * np.sin(yt), np.cos(xt) * np.cos(yt)], axis=1)
        d = local @ R.T
        src = t[None, :] + d * self.depthMinimum
        return np.concatenate([src, d], axis=1)

    def simulate(self, engine, q, link_world_id: Optional[int] = None) -> np.ndarray:
        """depth readings (measurementCount,); link_world_id: world id of the link the sensor rides on (ignored by its own rays)"""
        T = engine.fk_batch(np.asarray(q, dtype=np.float64)[None, :])[0] if self.link >= 0 else None
        rays = self.rays(T)
        ids, dist, _ = engine.raycast_batch(q, rays, None if link_world_id is None else [link_world_id])
        depth = np.where(ids >= 0, dist + self.depthMinimum, np.inf)
        return np.where((depth <= self.depthMinimum) | (depth >= self.depthMaximum), self.depthMaximum, depth)
