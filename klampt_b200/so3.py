"""Minimal so3 helpers in Klamp't's convention: a rotation is a 9-list in COLUMN-major order
[a11,a21,a31,a12,a22,a32,a13,a23,a33] (reference Python/klampt/math/so3.py:1-14), whereas the C ABI and the
file formats are ROW-major (Cpp/docs/Manual-FileTypes.md:35-37).  Only what the collision adapters need."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


def identity() -> List[float]:
    return [1., 0., 0., 0., 1., 0., 0., 0., 1.]


def matrix(R: Sequence[float]) -> np.ndarray:
    """column-major 9-list -> 3x3 array"""
    return np.asarray(R, dtype=np.float64).reshape(3, 3).T.copy()


def from_matrix(M) -> List[float]:
    return list(np.asarray(M, dtype=np.float64).T.reshape(-1))


def inv(R: Sequence[float]) -> List[float]:
    return from_matrix(matrix(R).T)


def mul(R1: Sequence[float], R2: Sequence[float]) -> List[float]:
    return from_matrix(matrix(R1) @ matrix(R2))


def apply(R: Sequence[float], p: Sequence[float]) -> List[float]:
    return list(matrix(R) @ np.asarray(p, dtype=np.float64))


def from_axis_angle(aa: Tuple[Sequence[float], float]) -> List[float]:
    axis, angle = aa
    w = np.asarray(axis, dtype=np.float64)
    w = w / np.linalg.norm(w)
    c, s = math.cos(angle), math.sin(angle)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return from_matrix(c * np.eye(3) + (1 - c) * np.outer(w, w) + s * K)


def from_quaternion(q: Sequence[float]) -> List[float]:
    """q = (w,x,y,z), not necessarily normalised"""
    w, x, y, z = [float(v) for v in q]
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    M = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return from_matrix(M)


def rpy(R: Sequence[float]) -> Tuple[float, float, float]:
    """roll, pitch, yaw with R = Rz(yaw) Ry(pitch) Rx(roll).  Ranges follow the reference
    (Python/klampt/math/so3.py:84-116): roll and yaw in [0, 2 pi) away from the singularity; at pitch = +-pi/2 the
    roll is fixed to 0 and yaw = -asin(m01), reflected to pi - yaw when cos(yaw) and m11 disagree in sign."""
    M = matrix(R)
    pitch = -math.asin(min(1.0, max(M[2, 0], -1.0)))
    cp = math.cos(pitch)
    two_pi = 2.0 * math.pi
    if abs(cp) > 1e-7:
        yaw = math.atan2(M[1, 0] / cp, M[0, 0] / cp) % two_pi
        roll = math.atan2(M[2, 1] / cp, M[2, 2] / cp) % two_pi
        return roll, pitch, yaw
    yaw = -math.asin(min(1.0, max(M[0, 1], -1.0)))
    sgn = lambda x: int(x > 0) - int(x < 0)
    if sgn(math.cos(yaw)) != sgn(M[1, 1]):
        yaw = math.pi - yaw
    return 0.0, pitch, yaw


def to_rowmajor12(R: Sequence[float], t: Sequence[float]) -> np.ndarray:
    """Klamp't (R column-major, t) -> the C ABI's 12 doubles (row-major R, then t)"""
    return np.concatenate([matrix(R).reshape(-1), np.asarray(t, dtype=np.float64)])


def from_rowmajor12(T12) -> Tuple[List[float], List[float]]:
    T12 = np.asarray(T12, dtype=np.float64)
    return from_matrix(T12[:9].reshape(3, 3)), list(T12[9:12])
