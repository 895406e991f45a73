"""Camera path and object pose schedule of a PEGASUS scene, emitted for MANY frames at once so that
several frames can be in flight per GPU.  Host side, float64, tiny.

Camera path  — PegasusSetup.create_camera_trajectory (/root/reference/src/gs/pegasus_setup.py:85-143)
               with interpolate_pose (/root/reference/src/utility/pose_interpolation.py:58-106):
               SLERP on the rotation + lerp on the translation between consecutive COLMAP poses.
Pose schedule — static_object_pose / dynamic_object_pose / update_object_pose (:160-226) in ABSOLUTE
               form: the reference applies R_delta = R_k R_{k-1}^T, t_delta = t_k - t_{k-1} to the
               already transformed cloud each frame; algebraically x_k = R_k (x_0 - m_0) + m_0 + t_k
               (SURVEY §8 a-1), which is what the pose kernel evaluates from the canonical cloud.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Mapping, Sequence, Tuple

import numpy as np

from .cameras import Camera, focal2fov
from .sh_rotation import quat_xyzw_to_rotation, rotation_to_quat_wxyz


# ---------------------------------------------------------------------------------------------
# pose interpolation (pose_interpolation.py:21-106)
# ---------------------------------------------------------------------------------------------
def pose_matrix_to_quat(pose: np.ndarray) -> np.ndarray:
    """4x4 pose -> (qx, qy, qz, qw, x, y, z) (pose_interpolation.py:21-28).  The quaternion's sign is
    irrelevant downstream: SLERP flips it onto the other's hemisphere."""
    assert pose.shape == (4, 4)
    w, x, y, z = rotation_to_quat_wxyz(np.asarray(pose[:3, :3], dtype=np.float64))
    return np.array([x, y, z, w, pose[0, 3], pose[1, 3], pose[2, 3]], dtype=np.float64)


def pose_quat_to_matrix(pose: np.ndarray) -> np.ndarray:
    """(qx, qy, qz, qw, x, y, z) -> 4x4; float32 like the reference's (pose_interpolation.py:31-41)."""
    assert pose.size == 7
    p = np.eye(4, dtype=np.float32)
    p[:3, :3] = quat_xyzw_to_rotation(pose[:4])
    p[:3, 3] = pose[4:]
    return p


def quaternion_slerp(q1, q2, alpha: float) -> np.ndarray:
    """pose_interpolation.py:58-84: shortest-arc SLERP, nlerp above dot 0.9995."""
    assert 0.0 <= alpha <= 1.0
    q1 = np.asarray(q1, dtype=np.float64)
    q2 = np.asarray(q2, dtype=np.float64)
    dot = float(q1.dot(q2))
    if dot < 0:
        q1, dot = -q1, -dot
    if dot > 0.9995:
        res = q1 + alpha * (q2 - q1)
        return res / np.linalg.norm(res)
    theta_0 = math.acos(dot)
    theta = theta_0 * alpha
    s2 = math.sin(theta) / math.sin(theta_0)
    s1 = math.cos(theta) - dot * s2
    return s1 * q1 + s2 * q2


def interpolate_pose(t: float, t1: float, pose1: np.ndarray, t2: float, pose2: np.ndarray) -> np.ndarray:
    """pose_interpolation.py:87-106."""
    if pose1.shape == (4, 4):
        pose1 = pose_matrix_to_quat(pose1)
    if pose2.shape == (4, 4):
        pose2 = pose_matrix_to_quat(pose2)
    assert t1 <= t <= t2
    r = (float(t) - float(t1)) / (float(t2) - float(t1))
    pos = pose1[4:] + r * (pose2[4:] - pose1[4:])
    rot = quaternion_slerp(pose1[:4], pose2[:4], r)
    return pose_quat_to_matrix(np.hstack((rot, pos)))


def qvec2rotmat(qvec) -> np.ndarray:
    """COLMAP (w, x, y, z) -> rotation matrix (GSP/scene/colmap_loader.py:43-53)."""
    w, x, y, z = (float(v) for v in qvec)
    return np.array([
        [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
        [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
        [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def camera_path_poses(extrinsics: Mapping, num_cameras: int = 5, num_interpolation_steps: int = 24,
                      mode: str = "random", rng=np.random) -> List[Tuple[np.ndarray, np.ndarray]]:
    """[(R, T)] of create_camera_trajectory (pegasus_setup.py:85-111,132-133): `extrinsics` maps COLMAP
    image ids to objects with .qvec (wxyz) and .tvec.  `rng` is numpy's global generator by default,
    as in the reference; draws happen in the reference's order (start frame, then the zoom factors).

    Reference quirk kept: in 'random+zoom' mode BOTH zoom factors scale pose1's translation — the second
    `if` (:110-111) says pose1 where pose2 was evidently meant."""
    if mode not in ("random", "sequence", "random+zoom"):
        raise ValueError(f"unknown mode {mode!r}")
    keys = sorted(extrinsics.keys())
    start = int(rng.randint(0, len(keys) - num_cameras))
    out = []
    for pose_idx in range(start, start + num_cameras):
        a, b = extrinsics[keys[pose_idx]], extrinsics[keys[pose_idx + 1]]
        pose1 = np.eye(4)
        pose1[:3, :3] = np.transpose(qvec2rotmat(a.qvec))
        pose1[:3, 3] = np.array(a.tvec)
        if mode == "random+zoom":
            pose1[:3, 3] *= rng.uniform(0.6, 1)
        pose2 = np.eye(4)
        pose2[:3, :3] = np.transpose(qvec2rotmat(b.qvec))
        pose2[:3, 3] = np.array(b.tvec)
        if mode == "random+zoom":
            pose1[:3, 3] *= rng.uniform(0.6, 1)
        for s in np.linspace(0, 1, num_interpolation_steps + 1)[:-1]:
            T = interpolate_pose(t=s, t1=0, pose1=pose1, t2=1, pose2=pose2)
            out.append((T[:3, :3], np.array(T[:3, 3])))
    return out


def create_camera_trajectory(extrinsics: Mapping, fx: float, image_width: int, image_height: int,
                             render_width: int, render_height: int, num_cameras: int = 5,
                             num_interpolation_steps: int = 24, mode: str = "random", rng=np.random,
                             device="cuda") -> List[Camera]:
    """Camera objects of create_camera_trajectory.  Reference quirks kept (pegasus_setup.py:119-122):
    fx is used for BOTH focal lengths, and the FoVs come from the ORIGINAL COLMAP image size while the
    image rendered has (render_width, render_height)."""
    fovy = focal2fov(float(fx), image_height)
    fovx = focal2fov(float(fx), image_width)
    return [Camera(R=R, T=T, FoVx=fovx, FoVy=fovy, image_width=render_width, image_height=render_height, device=device)
            for R, T in camera_path_poses(extrinsics, num_cameras, num_interpolation_steps, mode, rng)]


# ---------------------------------------------------------------------------------------------
# object pose schedule (pegasus_setup.py:160-226)
# ---------------------------------------------------------------------------------------------
def _step(trajectory: Mapping, object_id, step) -> Tuple[np.ndarray, np.ndarray]:
    e = trajectory[str(object_id)][str(step)]
    return quat_xyzw_to_rotation(np.asarray(e["q"], dtype=np.float64)), np.asarray(e["t"], dtype=np.float64)


def static_object_poses(trajectory: Mapping, object_ids: Iterable) -> List[Tuple[np.ndarray, np.ndarray]]:
    """static_object_pose (:208-226): every object takes the LAST step of body "1"'s key list
    (`list(self.object_trajectory[str(1)].keys())[-1]`, i.e. insertion order of the JSON)."""
    last = list(trajectory[str(1)].keys())[-1]
    return [_step(trajectory, oid, last) for oid in object_ids]


def dynamic_object_poses(trajectory: Mapping, object_ids: Sequence, num_frames: int,
                         first_step: int = 0) -> List[List[Tuple[np.ndarray, np.ndarray]]]:
    """Absolute (R_k, t_k) for frames 0..num_frames-1: dynamic_object_pose (:160-176) places step 0,
    update_object_pose (:178-196) then composes the per-step deltas, whose product telescopes to the
    trajectory's own (q_k, t_k).  frames[f][k] = pose of object_ids[k] at frame f."""
    return [[_step(trajectory, oid, first_step + f) for oid in object_ids] for f in range(num_frames)]


def pose_deltas(trajectory: Mapping, object_id, timestep: int) -> Tuple[np.ndarray, np.ndarray]:
    """(R_delta, t_delta) exactly as update_object_pose forms them (:180-189) — kept for parity tests
    of the absolute form against the reference's incremental one."""
    R1, t1 = _step(trajectory, object_id, timestep)
    R0, t0 = _step(trajectory, object_id, timestep - 1)
    return R1 @ R0.T, t1 - t0


def replay_recorded_drop(t: np.ndarray, q: np.ndarray, num_objects: int, num_frames: int, stride: int = 1,
                         stagger: int = 20, ring: float = 0.28) -> Dict[str, Dict[str, Dict[str, list]]]:
    """A K-object trajectory dict in the reference's JSON layout (`[body][step]{t, q}`, q as x,y,z,w —
    src/engine/physical_simulation.py / src/gs/pegasus_setup.py:160-196) made from ONE recorded body: object k
    replays the recording (every `stride`-th step) starting `k * stagger` steps later and shifted to a ring of radius
    `ring` metres, so K objects drop and tumble one after the other.  The recording is the reference's own
    src/engine/simulation_steps.json, body 1 (committed excerpt: tests/golden/simulation_body1.npz)."""
    t = np.asarray(t, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    n = t.shape[0]
    traj: Dict[str, Dict[str, Dict[str, list]]] = {}
    for k in range(num_objects):
        ang = 2.0 * np.pi * k / max(num_objects, 1)
        shift = np.array([ring * np.cos(ang), ring * np.sin(ang), 0.0]) - np.array([t[-1, 0], t[-1, 1], 0.0])
        body = {}
        for f in range(num_frames):
            i = min(max(f * stride - k * stagger, 0), n - 1)
            body[str(f)] = {"t": list(t[i] + shift), "q": list(q[i])}
        traj[str(k + 1)] = body
    return traj
