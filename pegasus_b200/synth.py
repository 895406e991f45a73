"""Seeded synthetic clouds, cameras and pose trajectories (the Zenodo datasets are not available
offline).  Distributions follow SURVEY.md §8(d): raw PLY-style parameters (logit opacity, log scale,
un-normalised quaternions), SH degree 3.  numpy only (host); nothing here is on the hot path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

from .cameras import focal2fov


def _sh_rest(rng, n):
    band = np.concatenate([np.full(3, 1 / 2), np.full(5, 1 / 3), np.full(7, 1 / 4)]).astype(np.float32)
    return (rng.normal(0, 0.15, size=(n, 15, 3)).astype(np.float32) * band[None, :, None])


def _opacity(rng, n):
    hi = rng.normal(3.0, 1.0, size=n)
    lo = rng.normal(-2.0, 1.5, size=n)
    return np.where(rng.random(n) < 0.6, hi, lo).astype(np.float32).reshape(n, 1)


def make_env(n: int, seed: int = 1000, extent: float = 1.5, scale_mu: float = 0.008) -> Dict[str, np.ndarray]:
    """Ground slab (70 %) + clutter boxes / walls (30 %) over [-extent, extent]^2."""
    rng = np.random.default_rng(seed)
    n_g = int(0.7 * n)
    n_c = n - n_g
    ground = np.stack([rng.uniform(-extent, extent, n_g), rng.uniform(-extent, extent, n_g),
                       rng.normal(0, 0.01, n_g)], axis=1)
    # clutter: points on the faces of a few boxes and two walls
    centers = rng.uniform(-extent * 0.8, extent * 0.8, size=(12, 2))
    sizes = rng.uniform(0.1, 0.4, size=(12, 3))
    which = rng.integers(0, 14, n_c)
    clutter = np.zeros((n_c, 3))
    for b in range(12):
        m = which == b
        k = int(m.sum())
        u = rng.uniform(-0.5, 0.5, size=(k, 3))
        face = rng.integers(0, 3, k)
        u[np.arange(k), face] = np.sign(u[np.arange(k), face]) * 0.5
        clutter[m] = np.concatenate([centers[b], [sizes[b, 2] / 2]]) + u * sizes[b]
    for wl, axis in ((12, 0), (13, 1)):
        m = which == wl
        k = int(m.sum())
        pts = np.stack([rng.uniform(-extent, extent, k), rng.uniform(-extent, extent, k), rng.uniform(0, 1.0, k)], axis=1)
        pts[:, axis] = extent + rng.normal(0, 0.005, k)
        clutter[m] = pts
    xyz = np.concatenate([ground, clutter]).astype(np.float32)
    perm = rng.permutation(n)  # trained clouds have no spatial order
    xyz = xyz[perm]
    scaling = np.clip(rng.normal(math.log(scale_mu), 0.6, size=(n, 3)), math.log(1e-4), math.log(0.2)).astype(np.float32)
    return dict(xyz=xyz, features_dc=rng.uniform(-1.7, 1.7, size=(n, 1, 3)).astype(np.float32),
                features_rest=_sh_rest(rng, n), opacity=_opacity(rng, n), scaling=scaling,
                rotation=rng.normal(size=(n, 4)).astype(np.float32))


def make_object(n: int, seed: int = 2000, kind: str = None) -> Dict[str, np.ndarray]:
    """Points on / near the surface of a 5-20 cm cylinder, box or ellipsoid centred near the origin."""
    rng = np.random.default_rng(seed)
    kind = kind or ("cylinder", "box", "ellipsoid")[seed % 3]
    dims = rng.uniform(0.05, 0.2, size=3) / 2
    if kind == "cylinder":
        th = rng.uniform(0, 2 * np.pi, n)
        side = rng.random(n) < 0.75
        r = np.where(side, 1.0, np.sqrt(rng.random(n)))
        z = np.where(side, rng.uniform(-1, 1, n), np.sign(rng.normal(size=n)))
        pts = np.stack([r * np.cos(th) * dims[0], r * np.sin(th) * dims[0], z * dims[2]], axis=1)
    elif kind == "box":
        u = rng.uniform(-1, 1, size=(n, 3))
        face = rng.integers(0, 3, n)
        u[np.arange(n), face] = np.sign(u[np.arange(n), face])
        pts = u * dims
    else:
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * dims
    pts = pts + rng.normal(0, 0.0008, size=(n, 3))
    scaling = np.clip(rng.normal(math.log(0.0015), 0.5, size=(n, 3)), math.log(1e-4), math.log(0.05)).astype(np.float32)
    return dict(xyz=pts.astype(np.float32), features_dc=rng.uniform(-1.7, 1.7, size=(n, 1, 3)).astype(np.float32),
                features_rest=_sh_rest(rng, n), opacity=_opacity(rng, n), scaling=scaling,
                rotation=rng.normal(size=(n, 4)).astype(np.float32))


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> Tuple[np.ndarray, np.ndarray]:
    """(R, T) in the reference Camera convention: R = camera-to-world rotation, T = W2C translation;
    camera looks along +z, y down."""
    eye, target = np.asarray(eye, float), np.asarray(target, float)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, float))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f], axis=1)
    T = -R.T @ eye
    return R, T


def orbit_cameras(n_views: int, W: int, H: int, seed: int = 3000, fovx_deg: float = 72.28,
                  radius=(0.5, 1.2), elevation_deg=(15.0, 65.0)) -> List[dict]:
    """Orbit views; fx = fy as the reference does (src/gs/pegasus_setup.py:119-122)."""
    rng = np.random.default_rng(seed)
    fovx = math.radians(fovx_deg)
    fovy = focal2fov(W / (2 * math.tan(fovx / 2)), H)
    cams = []
    for i in range(n_views):
        az = 2 * np.pi * i / max(n_views, 1) + rng.normal(0, 0.05)
        el = math.radians(rng.uniform(*elevation_deg))
        rad = rng.uniform(*radius)
        eye = np.array([rad * math.cos(el) * math.cos(az), rad * math.cos(el) * math.sin(az), rad * math.sin(el) + 0.05])
        tgt = rng.normal(0, 0.05, size=3) * np.array([1, 1, 0.3]) + np.array([0, 0, 0.05])
        R, T = look_at(eye, tgt)
        cams.append(dict(R=R, T=T, FoVx=fovx, FoVy=fovy, W=W, H=H))
    return cams


def random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def static_poses(k: int, seed: int = 4000, spread: float = 0.35) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Objects resting on the table plane with a random tumble."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(k):
        ang = 2 * np.pi * i / max(k, 1) + rng.normal(0, 0.2)
        rad = rng.uniform(0.05, spread)
        t = np.array([rad * math.cos(ang), rad * math.sin(ang), rng.uniform(0.04, 0.1)])
        out.append((random_rotation(rng), t))
    return out


def drop_trajectory(k: int, frames: int, seed: int = 4000) -> np.ndarray:
    """(frames, k, 7) [t(3), q xyzw(4)] — objects falling and tumbling to rest, the shape of a
    PyBullet recording (src/engine/physical_simulation.py:152), generated analytically because
    pybullet is not installed here."""
    rng = np.random.default_rng(seed)
    out = np.zeros((frames, k, 7))
    for i in range(k):
        ang = 2 * np.pi * i / max(k, 1)
        rad = rng.uniform(0.05, 0.35)
        x0, y0, z0 = rad * math.cos(ang), rad * math.sin(ang), rng.uniform(0.4, 0.7)
        rest = rng.uniform(0.04, 0.08)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        w0 = rng.uniform(4, 10)
        q0 = rng.normal(size=4)
        q0 /= np.linalg.norm(q0)
        t_hit = math.sqrt(2 * (z0 - rest) / 9.81)
        for f in range(frames):
            tt = f / 120.0
            if tt < t_hit:
                z = z0 - 0.5 * 9.81 * tt * tt
                ang_f = w0 * tt
            else:
                z = rest
                dt = tt - t_hit
                ang_f = w0 * t_hit + w0 * (1 - math.exp(-6 * dt)) / 6
            # rotation about `axis` by ang_f composed with q0 (xyzw)
            s, c = math.sin(ang_f / 2), math.cos(ang_f / 2)
            ax, ay, az_ = axis * s
            bx, by, bz, bw = q0
            q = np.array([c * bx + ax * bw + ay * bz - az_ * by,
                          c * by - ax * bz + ay * bw + az_ * bx,
                          c * bz + ax * by - ay * bx + az_ * bw,
                          c * bw - ax * bx - ay * by - az_ * bz])
            out[f, i, :3] = (x0, y0, z)
            out[f, i, 3:] = q / np.linalg.norm(q)
    return out
