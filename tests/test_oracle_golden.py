"""Pin the CPU oracle against outputs of the reference's own Python (tests/golden/, made by
tools/make_golden.py from /root/reference).  CPU only."""
import math

import numpy as np

import oracle


def test_sh_basis_matches_reference_eval_sh(golden):
    dirs, sh = golden["sh_dirs"], golden["sh_coeffs"]
    for deg in range(4):
        got = oracle.eval_sh(deg, sh, dirs)
        np.testing.assert_allclose(got, golden[f"sh_eval_deg{deg}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(oracle.rgb2sh(golden["rgb"]), golden["rgb2sh"], atol=1e-12)
    c = golden["sh_consts"]
    assert c[0] == oracle.SH_C0 and c[1] == oracle.SH_C1
    assert list(c[2:7]) == oracle.SH_C2 and list(c[7:14]) == oracle.SH_C3


def test_c_oracle_sh_to_rgb_matches_reference_eval_sh(golden):
    """The C preprocess colour (SH -> RGB, +0.5, clamp) against eval_sh from the reference."""
    dirs, sh = golden["sh_dirs"].astype(np.float32), golden["sh_coeffs"].astype(np.float32)
    N = dirs.shape[0]
    campos = np.array([0.3, -0.2, 0.1], np.float32)
    means = (campos + 2.5 * dirs).astype(np.float32)
    # camera at campos looking along +z of world: identity rotation; put everything in front by
    # using a synthetic view matrix that maps all points to depth 1 (only colour is checked)
    V = np.eye(4, dtype=np.float32); V[3, 2] = 0.0
    V = np.zeros((4, 4), np.float32); V[3, 2] = 1.0; V[3, 3] = 1.0   # p_view.z = 1 for every point
    M = np.zeros((4, 4), np.float32); M[3, 3] = 1.0                  # p_hom = (0,0,0,1) -> centre pixel
    shs = np.ascontiguousarray(sh.transpose(0, 2, 1))  # (N,16,3)
    for deg in range(4):
        pre = oracle.preprocess(means, np.ones(N, np.float32), V, M, campos, 64, 64, 1.0, 1.0, deg,
                                shs=shs, scales=np.full((N, 3), 0.01, np.float32),
                                rotations=np.tile(np.array([1, 0, 0, 0], np.float32), (N, 1)))
        assert (pre["radii"] > 0).all()
        d = (means - campos).astype(np.float64)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        want = np.maximum(oracle.eval_sh(deg, sh.astype(np.float64), d) + 0.5, 0.0)
        np.testing.assert_allclose(pre["rgb"], want, atol=3e-6)


def test_build_rotation_and_covariance(golden):
    q, s = golden["quat"], golden["scale"]
    np.testing.assert_allclose(oracle.build_rotation(q), golden["build_rotation"], atol=1e-6)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    for i in range(q.shape[0]):
        got = oracle.cov3d(s[i], 1.0, qn[i].astype(np.float32))
        np.testing.assert_allclose(got, golden["covariance"][i], rtol=2e-5, atol=1e-9)


def test_camera_matrices(golden):
    for i in range(golden["cam_R"].shape[0]):
        cam = oracle.camera(golden["cam_R"][i], golden["cam_T"][i], float(golden["cam_fovx"][i]),
                            float(golden["cam_fovy"][i]), 640, 480)
        np.testing.assert_array_equal(cam["world_view_transform"], golden["cam_wvt"][i])
        np.testing.assert_array_equal(cam["projection_matrix"], golden["cam_proj"][i])
        np.testing.assert_allclose(cam["full_proj_transform"], golden["cam_full"][i], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cam["camera_center"], golden["cam_center"][i], rtol=1e-5, atol=1e-6)


def test_generate_colors(golden):
    np.testing.assert_array_equal(oracle.generate_colors(7), golden["colors_bgr_7"])
    np.testing.assert_array_equal(oracle.generate_colors(5, "rgb"), golden["colors_rgb_5"])


def test_pose_xyz_and_quaternion(golden):
    xyz = oracle.apply_transformation_on_xyz(golden["pose_xyz_in"], golden["pose_R"], golden["pose_t"])
    np.testing.assert_allclose(xyz, golden["pose_xyz_out"], atol=1e-6)
    rot = oracle.apply_rotation_on_splats(golden["pose_rot_in"], golden["pose_R"])
    np.testing.assert_allclose(rot, golden["pose_rot_out"], atol=1e-6)


def test_merge_and_mask(golden):
    def cl(x):
        n = x.shape[0]
        return dict(xyz=x, features_dc=np.zeros((n, 1, 3), np.float32), features_rest=np.zeros((n, 15, 3), np.float32),
                    opacity=np.zeros((n, 1), np.float32), scaling=np.zeros((n, 3), np.float32),
                    rotation=np.zeros((n, 4), np.float32))
    m = oracle.merge_gaussians(cl(golden["merge_a_xyz"]), cl(golden["merge_b_xyz"]))
    np.testing.assert_array_equal(m["xyz"], golden["merge_xyz"])
    mask = np.ones(8, bool); mask[:5] = False
    np.testing.assert_array_equal(oracle.mask_points(m, mask)["xyz"], golden["masked_xyz"])


def test_pose_schedule(golden):
    steps = golden["traj_steps"]
    traj = {"1": {str(i): {"t": list(golden["traj_t"][i]), "q": list(golden["traj_q"][i])}
                  for i in range(len(steps))}}
    R, t = oracle.static_pose(traj, 1)
    T = golden["sched_static_T"]
    np.testing.assert_allclose(R, T[:3, :3], atol=1e-7)
    np.testing.assert_allclose(t, T[:3, 3], atol=1e-7)
    dyn = golden["sched_dynamic_T"]
    e0 = traj["1"]["0"]
    np.testing.assert_allclose(oracle.quat_xyzw_to_matrix(e0["q"]), dyn[0][:3, :3], atol=1e-6)
    for ts in range(1, 6):
        Rd, td = oracle.dynamic_pose_delta(traj, 1, ts)
        np.testing.assert_allclose(Rd, dyn[ts][:3, :3], atol=1e-6)
        np.testing.assert_allclose(td, dyn[ts][:3, 3], atol=1e-7)


def _mask_scene(golden):
    def cl(name):
        return dict(xyz=golden[f"mask_{name}_xyz"], features_dc=golden[f"mask_{name}_features_dc"],
                    features_rest=golden[f"mask_{name}_features_rest"], opacity=golden[f"mask_{name}_opacity"],
                    scaling=golden[f"mask_{name}_scaling"], rotation=golden[f"mask_{name}_rotation"])
    env = cl("env")
    objs = {int(k): cl(f"obj{int(k)}") for k in golden["mask_obj_order"]}
    W, H = [int(v) for v in golden["mask_WH"]]
    fovx, fovy = [float(v) for v in golden["mask_cam_fov"]]
    cam = oracle.camera(golden["mask_cam_R"], golden["mask_cam_T"], fovx, fovy, W, H)
    return cam, env, objs, golden["mask_colors"]


def test_k_plus_3_passes_match_reference_orchestration(golden):
    """src/gs/render.py's four helpers (run from the reference's text in tools/make_golden.py)
    against oracle.render_frame_reference."""
    cam, env, objs, colors = _mask_scene(golden)
    out = oracle.render_frame_reference(cam, env, objs, colors, np.zeros(3, np.float32))
    np.testing.assert_array_equal(out["rgb"], golden["mask_rgb"])
    np.testing.assert_array_equal(out["depth"], golden["mask_depth"])
    np.testing.assert_array_equal(out["silhouette"].astype(np.uint8), golden["mask_silhouette"])
    np.testing.assert_array_equal(out["visible"].astype(np.uint8), golden["mask_visible"])
    np.testing.assert_array_equal(out["sem_seg"], golden["mask_sem_seg"])
    assert golden["mask_silhouette"].sum() > 50 and golden["mask_visible"].sum() > 50


def test_expf_accuracy():
    xs = np.linspace(-20.0, 0.0, 20001).astype(np.float32)
    got = np.array([oracle.expf(float(x)) for x in xs], np.float64)
    ref = np.exp(xs.astype(np.float64))
    rel = np.abs(got - ref) / ref
    assert rel.max() < 1.6e-7, rel.max()
    assert oracle.expf(0.0) == 1.0 and oracle.expf(-100.0) == 0.0


def test_sh_rotation_invariance():
    """Y(d).(D c) == Y(R^T d).c — rotating the coefficients rotates the radiance field."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    R = Rotation.from_rotvec(rng.normal(size=3)).as_matrix()
    D1, D2, D3 = oracle.sh_rotation_matrices(R)
    for D in (D1, D2, D3):
        np.testing.assert_allclose(D @ D.T, np.eye(D.shape[0]), atol=1e-12)
    d = rng.normal(size=(50, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    c = rng.normal(size=(15,))
    cr = np.concatenate([D1 @ c[0:3], D2 @ c[3:8], D3 @ c[8:15]])
    lhs = oracle.sh_basis(d)[:, 1:] @ cr
    rhs = oracle.sh_basis(d @ R)[:, 1:] @ c
    np.testing.assert_allclose(lhs, rhs, atol=1e-12)
    # composition: D(R1 R2) = D(R1) D(R2)
    R2 = Rotation.from_rotvec(rng.normal(size=3)).as_matrix()
    for a, b, ab in zip(oracle.sh_rotation_matrices(R), oracle.sh_rotation_matrices(R2),
                        oracle.sh_rotation_matrices(R @ R2)):
        np.testing.assert_allclose(a @ b, ab, atol=1e-12)
