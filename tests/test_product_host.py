"""Host-side pieces of the PRODUCT (pegasus_b200/) pinned to the reference's own Python (golden vectors made by
tools/make_golden.py from /root/reference): the pose schedule (src/gs/pegasus_setup.py:160-226), the camera tensors
(GSP/scene/cameras.py:48-57, GSP/utils/graphics_utils.py:38-71) and the SH band rotations
(src/gs/gaussian_model.py:507-546).  The oracle has its own copies of these tests; these make sure the code the
GPU path is actually fed by says the same."""
import os

import numpy as np
import pytest

from pegasus_b200 import trajectory
from pegasus_b200.cameras import Camera
from pegasus_b200.sh_rotation import (band_values, pose_packet, quat_xyzw_to_rotation, rotation_to_quat_wxyz,
                                      sh_band_rotations)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden_traj(golden):
    n = len(golden["traj_steps"])
    return {"1": {str(i): {"t": list(golden["traj_t"][i]), "q": list(golden["traj_q"][i])} for i in range(n)}}


def test_static_pose_schedule_matches_reference(golden):
    """static_object_pose (pegasus_setup.py:208-226): the LAST step of body 1's key list."""
    (R, t), = trajectory.static_object_poses(_golden_traj(golden), [1])
    T = golden["sched_static_T"]
    np.testing.assert_allclose(R, T[:3, :3], atol=1e-7)
    np.testing.assert_allclose(t, T[:3, 3], atol=1e-7)


def test_dynamic_pose_schedule_matches_reference(golden):
    """dynamic_object_pose places step 0, update_object_pose composes (R_k R_{k-1}^T, t_k - t_{k-1}) per frame
    (pegasus_setup.py:160-196).  The product hands the kernels ABSOLUTE poses; their deltas must be the reference's,
    and composing the reference's deltas must give the absolute poses back."""
    traj = _golden_traj(golden)
    dyn = golden["sched_dynamic_T"]
    frames = trajectory.dynamic_object_poses(traj, [1], 6)
    R0, t0 = frames[0][0]
    np.testing.assert_allclose(R0, dyn[0][:3, :3], atol=1e-6)
    np.testing.assert_allclose(t0, dyn[0][:3, 3], atol=1e-7)
    R_acc, t_acc = dyn[0][:3, :3].astype(np.float64), dyn[0][:3, 3].astype(np.float64)
    for ts in range(1, 6):
        Rd, td = trajectory.pose_deltas(traj, 1, ts)
        np.testing.assert_allclose(Rd, dyn[ts][:3, :3], atol=1e-6)
        np.testing.assert_allclose(td, dyn[ts][:3, 3], atol=1e-7)
        R_acc, t_acc = dyn[ts][:3, :3].astype(np.float64) @ R_acc, t_acc + dyn[ts][:3, 3]
        Rk, tk = frames[ts][0]
        np.testing.assert_allclose(Rk, R_acc, atol=2e-6)   # float32 golden deltas, 5 products
        np.testing.assert_allclose(tk, t_acc, atol=1e-6)


def test_camera_tensors_match_reference(golden):
    """Camera == GSP Camera: transposed W2C, transposed projection, their product, centre = row 3 of the inverse."""
    for i in range(golden["cam_R"].shape[0]):
        cam = Camera(golden["cam_R"][i], golden["cam_T"][i], float(golden["cam_fovx"][i]), float(golden["cam_fovy"][i]),
                     640, 480, device="cpu")
        np.testing.assert_allclose(cam.world_view_transform.numpy(), golden["cam_wvt"][i], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cam.projection_matrix.numpy(), golden["cam_proj"][i], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cam.full_proj_transform.numpy(), golden["cam_full"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cam.camera_center.numpy(), golden["cam_center"][i], rtol=1e-5, atol=1e-6)
        # what the C ABI reads is the CONTIGUOUS transposed matrix: flat index [4 * col + row] of the W2C matrix
        assert cam.world_view_transform.is_contiguous() and cam.full_proj_transform.is_contiguous()


def _rot(seed):
    r = np.random.default_rng(seed)
    q = r.normal(size=4)
    return quat_xyzw_to_rotation(q)


def test_sh_band_1_closed_form():
    """SURVEY a-3: with e3nn's wigner_D(1, alpha, -beta, gamma) on the YXY angles of P^-1 R P, the l = 1 block the
    reference applies is D_1[i][j] = s_i s_j R[p_i][p_j], p = (1, 2, 0), s = (-1, +1, -1) — the 3DGS basis is
    (-C1 y, C1 z, -C1 x).  The product builds D by solving Y(S) D = Y(S R); the two must agree."""
    p, s = (1, 2, 0), (-1.0, 1.0, -1.0)
    for seed in range(8):
        R = _rot(seed)
        D1 = sh_band_rotations(R)[0]
        closed = np.array([[s[i] * s[j] * R[p[i], p[j]] for j in range(3)] for i in range(3)])
        np.testing.assert_allclose(D1, closed, atol=1e-12)


def test_sh_band_rotations_are_a_representation():
    """D_l(R) is orthogonal, D_l(I) = I, D_l(R1 R2) = D_l(R1) D_l(R2), and rotating the coefficients rotates the
    radiance: Y_l(d) . (D_l c) = Y_l(R^T d) . c for every direction d (gaussian_model.py:507-546's intent)."""
    r = np.random.default_rng(5)
    d = r.normal(size=(64, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    R1, R2 = _rot(11), _rot(12)
    for l, (Da, Db, Dab, Di) in enumerate(zip(sh_band_rotations(R1), sh_band_rotations(R2), sh_band_rotations(R1 @ R2),
                                              sh_band_rotations(np.eye(3)))):
        n = 2 * (l + 1) + 1
        np.testing.assert_allclose(Di, np.eye(n), atol=1e-12)
        np.testing.assert_allclose(Da @ Da.T, np.eye(n), atol=1e-12)
        np.testing.assert_allclose(Dab, Da @ Db, atol=1e-12)
        c = r.normal(size=n)
        lhs = band_values(d)[l] @ (Da @ c)
        rhs = band_values(d @ R1)[l] @ c   # rows of d @ R1 are R1^T d
        np.testing.assert_allclose(lhs, rhs, atol=1e-12)


def test_quaternion_round_trip_and_packet_layout():
    for seed in range(6):
        R = _rot(20 + seed)
        w, x, y, z = rotation_to_quat_wxyz(R)
        np.testing.assert_allclose(quat_xyzw_to_rotation([x, y, z, w]), R, atol=1e-12)
    pk = pose_packet(_rot(3), [0.1, 0.2, 0.3], [1.0, 2.0, 3.0], rotate_sh=True)
    assert pk.shape == (103,) and pk.dtype == np.float32 and pk[102:103].view(np.int32)[0] == 1
    np.testing.assert_allclose(pk[9:12], [0.1, 0.2, 0.3], atol=1e-7)
    np.testing.assert_allclose(pk[12:15], [1.0, 2.0, 3.0], atol=1e-7)


def test_recorded_drop_replay_uses_the_reference_recording():
    """tests/golden/simulation_body1.npz is body 1 of the reference's src/engine/simulation_steps.json
    (tools/make_golden_traj.py).  replay_recorded_drop lays K staggered copies out in the reference's JSON layout, so
    the reference's schedule functions run on it unchanged."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "simulation_body1.npz"))
    assert g["t"].shape == (4000, 3) and g["q"].shape == (4000, 4)
    assert 0.5 < g["t"][0, 2] < 0.6 and 0.05 < g["t"][-1, 2] < 0.07          # dropped from 0.55 m, rests at 6 cm
    np.testing.assert_allclose(np.linalg.norm(g["q"], axis=1), 1.0, atol=1e-6)
    traj = trajectory.replay_recorded_drop(g["t"], g["q"], num_objects=3, num_frames=50, stride=1, stagger=10)
    assert sorted(traj.keys()) == ["1", "2", "3"] and len(traj["1"]) == 50
    frames = trajectory.dynamic_object_poses(traj, [1, 2, 3], 50)
    # object 1 follows the recording step by step (shifted in xy only); object 2 starts 10 frames later
    np.testing.assert_allclose(frames[7][0][0], quat_xyzw_to_rotation(g["q"][7]), atol=1e-12)
    np.testing.assert_allclose(frames[7][0][1][2], g["t"][7, 2], atol=1e-12)
    np.testing.assert_allclose(frames[17][1][0], quat_xyzw_to_rotation(g["q"][7]), atol=1e-12)
    # the reference's incremental deltas telescope to the absolute poses the kernels get
    R, t = frames[0][2]
    for ts in range(1, 50):
        Rd, td = trajectory.pose_deltas(traj, 3, ts)
        R, t = Rd @ R, t + td
    np.testing.assert_allclose(R, frames[49][2][0], atol=1e-9)
    np.testing.assert_allclose(t, frames[49][2][1], atol=1e-12)
