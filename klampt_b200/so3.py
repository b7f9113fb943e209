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
    roll is fixed to 0 and yaw = -asin(m01), reflected to pi - yaw when cos(yaw) and m11 disagree in sign.

    One stated deviation: the reference takes acos of the cosine and returns 2 pi - acos(..) whenever the sine's sign differs from
    cos(pitch)'s, a zero sine included, so it answers (2 pi, 0, 2 pi) for the identity and 2 pi wherever m10 or m21 is exactly 0.  This
    one uses atan2 (better conditioned near 0 and pi) folded into [0, 2 pi), so those cases answer 0.  The rotation is the same;
    only the representative of the angle modulo 2 pi differs (tests/test_oracle_kat.py pins both facts)."""
    M = matrix(R)
    pitch = -math.asin(min(1.0, max(M[2, 0], -1.0)))
    cp = math.cos(pitch)
    two_pi = 2.0 * math.pi
    if abs(cp) > 1e-7:
        fold = lambda a: 0.0 if (a % two_pi) >= two_pi else a % two_pi      # -1e-17 % 2 pi rounds up to 2 pi itself
        yaw = fold(math.atan2(M[1, 0] / cp, M[0, 0] / cp))
        roll = fold(math.atan2(M[2, 1] / cp, M[2, 2] / cp))
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


# ---- Euler ZYX triplets of Floating / BallAndSocket joints (row-major 3x3 numpy arrays here, not Klamp't 9-lists) ----------
def euler_zyx_matrix(a: float, b: float, c: float) -> np.ndarray:
    """EulerAngleRotation(a,b,c).getMatrixZYX = Rz(a) Ry(b) Rx(c) (the three links turn about z, y, x;
    reference Cpp/Modeling/Interpolate.cpp:24-30)."""
    ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
    return np.array([[ca * cb, ca * sb * sc - sa * cc, ca * sb * cc + sa * sc],
                     [sa * cb, sa * sb * sc + ca * cc, sa * sb * cc - ca * sc],
                     [-sb, cb * sc, cb * cc]])


def matrix_euler_zyx(R: np.ndarray) -> Tuple[float, float, float]:
    b = math.asin(min(1.0, max(-1.0, -R[2, 0])))
    if abs(R[2, 0]) < 1.0 - 1e-12:
        return math.atan2(R[1, 0], R[0, 0]), b, math.atan2(R[2, 1], R[2, 2])
    return math.atan2(-R[0, 1], R[1, 1]), b, 0.0      # gimbal lock: the whole turn goes into the z angle


def angle(R: np.ndarray) -> float:
    """AngleAxisRotation::angle of a rotation matrix"""
    return math.acos(min(1.0, max(-1.0, 0.5 * (np.trace(R) - 1.0))))


def log(R: np.ndarray) -> np.ndarray:
    th = angle(R)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        return 0.5 * v
    if math.pi - th < 1e-6:      # near a half turn: axis from the diagonal of (R + I) / 2
        ax = np.sqrt(np.maximum(0.5 * (np.diag(R) + 1.0), 0.0))
        m = int(np.argmax(ax))
        for k in range(3):
            if k != m and R[m, k] + R[k, m] < 0:
                ax[k] = -ax[k]
        if ax @ v < 0:
            ax = -ax
        return th * ax / np.linalg.norm(ax)
    return th / (2.0 * math.sin(th)) * v


def exp(w: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    x = np.asarray(w) / th
    c, s = math.cos(th), math.sin(th)
    return c * np.eye(3) + (1 - c) * np.outer(x, x) + s * K / th


def euler_zyx_interp(ea, eb, u: float) -> Tuple[float, float, float]:
    """interpolateRotation on two Euler ZYX triplets: Ra exp(u log(Ra^T Rb)), back to a triplet"""
    Ra, Rb = euler_zyx_matrix(*ea), euler_zyx_matrix(*eb)
    return matrix_euler_zyx(Ra @ exp(u * log(Ra.T @ Rb)))


def euler_zyx_angle_between(ea, eb) -> float:
    return angle(euler_zyx_matrix(*ea) @ euler_zyx_matrix(*eb).T)
