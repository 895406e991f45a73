"""Host-side pose packet maths: rotation -> quaternion and real-SH band rotation matrices.

Replaces the e3nn / scipy round trips of GaussianModel.apply_rotation_on_sh and
apply_rotation_on_splats (/root/reference/src/gs/gaussian_model.py:499-546).  D_l is the matrix
with  Y_l(d) . (D_l c) = Y_l(R^T d) . c  in the 3DGS real-SH basis
(/root/reference/submodules/gaussian-splatting-pegasus/utils/sh_utils.py:74-100), i.e. rotating the
coefficients rotates the radiance field with the object.  It is obtained exactly (float64) by
evaluating the basis on a fixed direction set S and solving  Y(S) D = Y(S R):
D = pinv(Y(S)) @ Y(S R); pinv(Y(S)) is computed once at import.
"""
from __future__ import annotations

import numpy as np

C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def band_values(d: np.ndarray):
    """d: (N,3) unit vectors -> (Y1 (N,3), Y2 (N,5), Y3 (N,7)) in coefficient order."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xx, yy, zz = x * x, y * y, z * z
    Y1 = np.stack([-C1 * y, C1 * z, -C1 * x], axis=1)
    Y2 = np.stack([C2[0] * x * y, C2[1] * y * z, C2[2] * (2 * zz - xx - yy), C2[3] * x * z, C2[4] * (xx - yy)], axis=1)
    Y3 = np.stack([C3[0] * y * (3 * xx - yy), C3[1] * x * y * z, C3[2] * y * (4 * zz - xx - yy),
                   C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                   C3[6] * x * (xx - 3 * yy)], axis=1)
    return Y1, Y2, Y3


def _fibonacci_sphere(n: int) -> np.ndarray:
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


_S = _fibonacci_sphere(48)
_PINV = [np.linalg.pinv(Y) for Y in band_values(_S)]


def sh_band_rotations(R: np.ndarray):
    """(D1 3x3, D2 5x5, D3 7x7), float64."""
    R = np.asarray(R, dtype=np.float64)
    rotated = band_values(_S @ R)  # rows are R^T s
    return [P @ Y for P, Y in zip(_PINV, rotated)]


def rotation_to_quat_wxyz(R: np.ndarray) -> np.ndarray:
    """Unit quaternion (w,x,y,z) of a rotation matrix (Shepperd's method), float64."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def quat_xyzw_to_rotation(q) -> np.ndarray:
    """Rotation matrix of a (x,y,z,w) quaternion — the order PyBullet / the trajectory JSON uses
    (/root/reference/src/gs/pegasus_setup.py:165-169)."""
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


POSE_WORDS = 103


def pose_packet(R, t, pivot, rotate_sh: bool = True) -> np.ndarray:
    """One pg_pose as 103 float32 words (the last word is the int32 rotate_sh flag)."""
    R = np.asarray(R, dtype=np.float64)
    D1, D2, D3 = sh_band_rotations(R)
    out = np.zeros(POSE_WORDS, dtype=np.float32)
    out[0:9] = R.reshape(-1)
    out[9:12] = np.asarray(t, dtype=np.float64)
    out[12:15] = np.asarray(pivot, dtype=np.float64)
    out[15:19] = rotation_to_quat_wxyz(R)
    out[19:28] = D1.reshape(-1)
    out[28:53] = D2.reshape(-1)
    out[53:102] = D3.reshape(-1)
    out[102:103].view(np.int32)[0] = 1 if rotate_sh else 0
    return out


def generate_pose_packets(poses, pivots, rotate_sh: bool = True) -> np.ndarray:
    """(K,103) float32 packet array for K absolute poses [(R, t), ...] and per-object pivots."""
    return np.stack([pose_packet(R, t, pv, rotate_sh) for (R, t), pv in zip(poses, pivots)])
