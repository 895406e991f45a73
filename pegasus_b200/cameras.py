"""Camera tensors with the reference's semantics
(/root/reference/submodules/gaussian-splatting-pegasus/scene/cameras.py:48-57 and
utils/graphics_utils.py:38-77): world_view_transform and full_proj_transform are stored TRANSPOSED,
znear 0.01 / zfar 100, camera_center is row 3 of the inverse view transform."""
from __future__ import annotations

import math

import numpy as np
import torch


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def world_to_view(R, t, translate=(0.0, 0.0, 0.0), scale=1.0) -> np.ndarray:
    """W2C 4x4 (float32) from the camera-to-world rotation R (COLMAP R transposed) and W2C translation t."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = np.asarray(R, dtype=np.float64).T
    Rt[:3, 3] = np.asarray(t, dtype=np.float64)
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    c2w[:3, 3] = (c2w[:3, 3] + np.asarray(translate, dtype=np.float64)) * scale
    return np.linalg.inv(c2w).astype(np.float32)


def projection(znear, zfar, fovX, fovY) -> torch.Tensor:
    ty, tx = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class Camera:
    def __init__(self, R, T, FoVx, FoVy, image_width, image_height, device="cuda", znear=0.01, zfar=100.0,
                 trans=(0.0, 0.0, 0.0), scale=1.0):
        self.R, self.T = np.asarray(R), np.asarray(T)
        self.FoVx, self.FoVy = float(FoVx), float(FoVy)
        self.image_width, self.image_height = int(image_width), int(image_height)
        self.znear, self.zfar = znear, zfar
        # matrix products / inverse on the host (tiny) so that no device sync is needed later
        wvt = torch.tensor(world_to_view(R, T, trans, scale)).transpose(0, 1)
        proj = projection(znear, zfar, self.FoVx, self.FoVy).transpose(0, 1)
        full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
        center = wvt.inverse()[3, :3]
        self.world_view_transform = wvt.contiguous().to(device)
        self.projection_matrix = proj.contiguous().to(device)
        self.full_proj_transform = full.contiguous().to(device)
        self.camera_center = center.contiguous().to(device)

    def intrinsics(self) -> np.ndarray:
        fx = fov2focal(self.FoVx, self.image_width)
        fy = fov2focal(self.FoVy, self.image_height)
        return np.array([[fx, 0, self.image_width / 2], [0, fy, self.image_height / 2], [0, 0, 1]])
