"""GPU parity tests: libpegasus_b200.so (through the C ABI / drop-in API) against the CPU oracle on
identical inputs.  Bars (BASELINE.json north_star): radii, tile keys, ranges, sorted order
bit-exact; RGB <= 1e-3 max-abs; depth <= 1e-4 relative; masks exact outside the 1e-5 threshold band.
Because the kernels and the oracle share an explicit IEEE operation order (and a software exp), the
images are additionally expected to be bit-identical, which is asserted where noted.
"""
import math

import numpy as np
import pytest
import torch

import oracle
from tests import util

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-3
DEPTH_RTOL = 1e-4


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def assert_fast_close(got, ref, what, tol, rel=False, budget=2e-5, cap=5e-2):
    """FAST numerics (MUFU exp, one blend weight per Gaussian) against the oracle: within `tol` everywhere except at
    the few pixels where a transmittance lands on the other side of the 1e-4 termination threshold (one Gaussian more
    or less blended — the same pixels differ between the reference's own MUFU-based expf and any other exp); those
    are budgeted (at most max(3, 2e-5 of the pixels)) and capped."""
    err = np.abs(got - ref)
    if rel:
        err = err / np.maximum(np.abs(ref), 1e-6)
    n_bad = int((err > tol).sum())
    assert n_bad <= max(3, int(budget * err.size)), f"{what}: {n_bad} of {err.size} values differ by more than {tol}"
    assert float(err.max()) <= cap, f"{what}: max difference {float(err.max())}"


def gpu_forward(inp, ocam, bg, sh_degree=3, reference_lists=False, numerics=None, **kw):
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    d = dev()
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(d)
    settings = GaussianRasterizationSettings(
        image_height=ocam["image_height"], image_width=ocam["image_width"],
        tanfovx=math.tan(ocam["FoVx"] * 0.5), tanfovy=math.tan(ocam["FoVy"] * 0.5), bg=t(np.asarray(bg, np.float32)),
        scale_modifier=kw.get("scale_modifier", 1.0), viewmatrix=t(ocam["world_view_transform"]),
        projmatrix=t(ocam["full_proj_transform"]), sh_degree=sh_degree, campos=t(ocam["camera_center"]),
        prefiltered=False, debug=kw.get("debug", False))
    rast = GaussianRasterizer(raster_settings=settings)
    rast.reference_lists = reference_lists
    rast.numerics = numerics
    means3D = t(inp["means3D"])
    color, radii, depth = rast(
        means3D=means3D, means2D=torch.zeros_like(means3D), shs=t(kw.get("shs", inp.get("shs"))),
        colors_precomp=t(kw.get("colors_precomp")), opacities=t(inp["opacities"]),
        scales=t(kw.get("scales", inp.get("scales"))), rotations=t(kw.get("rotations", inp.get("rotations"))),
        cov3D_precomp=t(kw.get("cov3D_precomp")))
    return color, radii, depth, rast.aux


def assert_sublists(ranges_sub, plist_sub, ranges_ref, plist_ref):
    """Per tile, the stored list must be a subsequence (same order) of the reference's list."""
    for t in range(ranges_ref.shape[0]):
        a = plist_sub[ranges_sub[t, 0]:ranges_sub[t, 1]]
        b = plist_ref[ranges_ref[t, 0]:ranges_ref[t, 1]]
        if a.size == 0:
            continue
        assert a.size <= b.size
        # greedy subsequence match: Gaussian ids are unique inside a tile list
        pos = {int(v): i for i, v in enumerate(b)}
        idx = np.fromiter((pos.get(int(v), -1) for v in a), dtype=np.int64, count=a.size)
        assert (idx >= 0).all() and (np.diff(idx) > 0).all(), f"tile {t}: stored list is not a subsequence"


def check_against_oracle(inp, ocam, bg, sh_degree=3, expect_bitexact_images=True, **kw):
    """Two GPU runs against one oracle run: (1) reference_lists=True — the binning state (64-bit keys,
    point list, ranges, n_contrib) must equal the reference's bit for bit; (2) the default path, which
    stores only pairs that can contribute — identical radii / R / images / final_T, and per tile a
    subsequence of the reference's list."""
    from pegasus_b200.scene import export_binning
    ref = util.oracle_forward(inp, ocam, bg, sh_degree, **kw)
    W, H = ocam["image_width"], ocam["image_height"]
    P = inp["means3D"].shape[0]
    denom = np.maximum(np.abs(ref["depth"]), 1e-6)
    full = None
    for reference_lists in (True, False):
        color, radii, depth, aux = gpu_forward(inp, ocam, bg, sh_degree, reference_lists=reference_lists, **kw)
        np.testing.assert_array_equal(radii.cpu().numpy(), ref["radii"])
        assert aux["num_rendered"] == ref["num_rendered"]
        assert aux["num_visible"] == int((ref["radii"] > 0).sum())
        keys, plist, ranges = export_binning(dev(), P, W, H, aux["pair_capacity"], aux["num_stored"])
        if reference_lists:
            assert aux["num_stored"] == ref["num_rendered"]
            np.testing.assert_array_equal(ranges, ref["ranges"])
            np.testing.assert_array_equal(keys, ref["keys"])
            np.testing.assert_array_equal(plist, ref["point_list"])
            np.testing.assert_array_equal(aux["n_contrib"].cpu().numpy().view(np.uint32), ref["n_contrib"])
            full = (color, radii, depth, aux)
        else:
            assert aux["num_stored"] <= ref["num_rendered"] and "n_contrib" not in aux
            assert bool((keys[1:] >= keys[:-1]).all())
            if ref["num_rendered"] <= 3_000_000:
                assert_sublists(ranges, plist, ref["ranges"], ref["point_list"])
        c, d = color.cpu().numpy(), depth.cpu().numpy()
        assert np.abs(c - ref["color"]).max() <= RGB_TOL
        assert (np.abs(d - ref["depth"]) / denom).max() <= DEPTH_RTOL
        np.testing.assert_array_equal(aux["final_T"].cpu().numpy(), ref["final_T"])
        if expect_bitexact_images:
            np.testing.assert_array_equal(c, ref["color"])
            np.testing.assert_array_equal(d, ref["depth"])
    # the product's default numerics: same integers, images inside the reference tolerances
    color, radii, depth, aux = gpu_forward(inp, ocam, bg, sh_degree, numerics="fast", **kw)
    np.testing.assert_array_equal(radii.cpu().numpy(), ref["radii"])
    assert aux["num_rendered"] == ref["num_rendered"]
    assert_fast_close(color.cpu().numpy(), ref["color"], "fast RGB", RGB_TOL)
    assert_fast_close(depth.cpu().numpy(), ref["depth"], "fast depth", DEPTH_RTOL, rel=True)
    assert_fast_close(aux["final_T"].cpu().numpy(), ref["final_T"], "fast final_T", 1e-5, cap=1e-2)
    return ref, full


def test_forward_bitexact_small_640x480():
    env, objs = util.small_scene()
    cloud = util.merged(env, objs)
    inp = util.activated(cloud)
    from pegasus_b200 import synth
    for ci, c in enumerate(synth.orbit_cameras(3, 640, 480, seed=3000)):
        ref, _ = check_against_oracle(inp, util.oracle_cam(c), np.zeros(3, np.float32))
        assert ref["num_rendered"] > 10000 and ref["color"].max() > 0.1


def test_forward_ragged_image_and_white_bg():
    env, objs = util.small_scene(n_env=6000, n_obj=(1500,), seed=3)
    inp = util.activated(util.merged(env, objs))
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 333, 205, seed=11)[0]   # not a multiple of 16 in either direction
    check_against_oracle(inp, util.oracle_cam(c), np.array([1.0, 1.0, 1.0], np.float32))
    check_against_oracle(inp, util.oracle_cam(c), np.array([0.2, 0.5, 0.9], np.float32), scale_modifier=0.6)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_forward_sh_degrees(deg):
    env, objs = util.small_scene(n_env=4000, n_obj=(1000,), seed=5)
    inp = util.activated(util.merged(env, objs))
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 320, 240, seed=12)[0]
    check_against_oracle(inp, util.oracle_cam(c), np.zeros(3, np.float32), sh_degree=deg)


def test_forward_precomputed_colors_and_covariance():
    """The reference's own alternative paths (GSP/gaussian_renderer/__init__.py:64-66,75-80)."""
    env, objs = util.small_scene(n_env=5000, n_obj=(1200,), seed=7)
    inp = util.activated(util.merged(env, objs))
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 320, 240, seed=13)[0]
    ocam = util.oracle_cam(c)
    P = inp["means3D"].shape[0]
    # convert_SHs_python branch: colours from eval_sh (float64 here) -> colours within 1e-5 of the SH path
    d = inp["means3D"].astype(np.float64) - ocam["camera_center"].astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cols = np.maximum(oracle.eval_sh(3, inp["shs"].transpose(0, 2, 1), d) + 0.5, 0).astype(np.float32)
    ref_sh = util.oracle_forward(inp, ocam, np.zeros(3, np.float32))
    ref, got = check_against_oracle(inp, ocam, np.zeros(3, np.float32), shs=None, colors_precomp=cols)
    assert np.abs(ref["color"] - ref_sh["color"]).max() < 1e-4
    # compute_cov3D_python branch
    cov = np.stack([oracle.cov3d(inp["scales"][i], 1.0, inp["rotations"][i]) for i in range(P)])
    ref2, _ = check_against_oracle(inp, ocam, np.zeros(3, np.float32), scales=None, rotations=None, cov3D_precomp=cov)
    np.testing.assert_array_equal(ref2["radii"], ref_sh["radii"])


def test_edge_cases_empty_culled_single_and_huge():
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 160, 96, seed=14)[0]
    ocam = util.oracle_cam(c)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    # P == 0: upstream returns zero images (the kernels never run), not the background
    empty = dict(means3D=np.zeros((0, 3), np.float32), opacities=np.zeros((0, 1), np.float32),
                 scales=np.zeros((0, 3), np.float32), rotations=np.zeros((0, 4), np.float32),
                 shs=np.zeros((0, 16, 3), np.float32))
    color, radii, depth, aux = gpu_forward(empty, ocam, bg)
    assert radii.numel() == 0 and float(color.abs().max()) == 0.0 and float(depth.abs().max()) == 0.0
    # everything behind the camera: all culled -> background everywhere, radii 0
    env, _ = util.small_scene(n_env=500, n_obj=(), seed=9)
    inp = util.activated(env)
    center = -c["R"] @ c["T"]                      # camera centre; it looks along +R[:,2]
    inp["means3D"] = (inp["means3D"] * 0.01 + center - 3.0 * c["R"][:, 2]).astype(np.float32)
    ref, got = check_against_oracle(inp, ocam, bg)
    assert ref["num_rendered"] == 0 and (ref["radii"] == 0).all()
    np.testing.assert_array_equal(got[0].cpu().numpy(), np.broadcast_to(bg[:, None, None], (3, 96, 160)))
    # a single Gaussian, and one so large that it covers every tile
    one = dict(means3D=np.array([[0.0, 0.0, 0.05]], np.float32), opacities=np.array([[0.9]], np.float32),
               scales=np.array([[0.01, 0.02, 0.01]], np.float32), rotations=np.array([[1, 0, 0, 0]], np.float32),
               shs=np.random.default_rng(0).normal(size=(1, 16, 3)).astype(np.float32))
    check_against_oracle(one, ocam, bg)
    huge = dict(one)
    huge["scales"] = np.array([[5.0, 5.0, 5.0]], np.float32)
    ref, _ = check_against_oracle(huge, ocam, bg)
    assert ref["num_rendered"] == 10 * 6


def test_binning_tall_rectangles_span_several_row_windows():
    """The binning kernel flattens rectangles into tile rows and processes them in windows of 3584 rows
    per 512-Gaussian chunk: big splats (tens of rows each) force several windows per chunk, in both the
    reference-list and the culled mode."""
    env, _ = util.small_scene(n_env=3000, n_obj=(), seed=31)
    env["scaling"] = env["scaling"] + 2.5   # ~12x larger: most rectangles cover a large part of the 30 tile rows
    inp = util.activated(env)
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 640, 480, seed=16)[0]
    ref, _ = check_against_oracle(inp, util.oracle_cam(c), np.zeros(3, np.float32))
    vis = ref["radii"] > 0
    assert ref["num_rendered"] / max(int(vis.sum()), 1) > 150   # > 150 tiles per visible Gaussian on average


def test_workspace_overflow_is_detected_and_retried():
    from pegasus_b200 import rasterizer
    env, objs = util.small_scene(n_env=8000, n_obj=(), seed=21)
    env["scaling"] = env["scaling"] + 1.5   # bigger splats: many pairs per Gaussian
    inp = util.activated(env)
    from pegasus_b200 import synth
    c = synth.orbit_cameras(1, 640, 480, seed=15)[0]
    old = rasterizer.default_pair_capacity
    try:
        rasterizer.default_pair_capacity = lambda P, W, H: 4096   # far too small on purpose
        ref, got = check_against_oracle(inp, util.oracle_cam(c), np.zeros(3, np.float32))
        assert ref["num_rendered"] > 4096 and got[3]["pair_capacity"] >= ref["num_rendered"]
    finally:
        rasterizer.default_pair_capacity = old
        rasterizer._PAIR_CAPACITY_HINT.clear()


def test_non_cuda_inputs_fail_loudly():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                      torch.zeros(3), False, False)
    with pytest.raises(RuntimeError):
        GaussianRasterizer(s)(means3D=torch.zeros(1, 3), means2D=torch.zeros(1, 3), opacities=torch.ones(1, 1),
                              shs=torch.zeros(1, 16, 3), scales=torch.ones(1, 3), rotations=torch.ones(1, 4))
    with pytest.raises(Exception):
        GaussianRasterizer(s)(means3D=torch.zeros(1, 3, device="cuda"), means2D=None, opacities=torch.ones(1, 1))


def test_mark_visible():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    env, _ = util.small_scene(n_env=3000, n_obj=(), seed=2)
    from pegasus_b200 import synth
    c = util.oracle_cam(synth.orbit_cameras(1, 64, 64, seed=1)[0])
    d = dev()
    s = GaussianRasterizationSettings(64, 64, 1.0, 1.0, torch.zeros(3, device=d), 1.0,
                                      torch.from_numpy(c["world_view_transform"]).to(d),
                                      torch.from_numpy(c["full_proj_transform"]).to(d), 3,
                                      torch.from_numpy(c["camera_center"]).to(d), False, False)
    vis = GaussianRasterizer(s).markVisible(torch.from_numpy(env["xyz"]).to(d)).cpu().numpy()
    p = env["xyz"]
    zview = (p @ c["world_view_transform"][:3, 2] + c["world_view_transform"][3, 2])
    sure = np.abs(zview - 0.2) > 1e-4
    np.testing.assert_array_equal(vis[sure], (zview > 0.2)[sure])


# ------------------------------------------------------------------------------------------------
# pose kernel (row a-1..a-3) and composed scene (a-5)
# ------------------------------------------------------------------------------------------------
def _pose_list(k, seed):
    from pegasus_b200 import synth
    return synth.static_poses(k, seed=seed)


@pytest.mark.parametrize("sh_mode", ["rotate", "canonical"])
def test_pose_kernel_matches_reference_transform(sh_mode):
    from pegasus_b200 import ComposedScene
    env, objs = util.small_scene(n_env=2000, n_obj=(1501, 777, 1024), seed=31)   # ragged sizes: unaligned SH blocks
    colors = oracle.generate_colors(3)
    sc = ComposedScene(env, objs, colors, sh_mode=sh_mode)
    poses = _pose_list(3, 5)
    sc.set_poses(poses)
    torch.cuda.synchronize()
    lo = sc.n_env
    for k, oid in enumerate(sc.object_ids):
        R, t = poses[k]
        want = oracle.apply_transformation(objs[oid], R.astype(np.float32), t.astype(np.float32), sh_mode=sh_mode)
        n = objs[oid]["xyz"].shape[0]
        got_xyz = sc.means3D[lo:lo + n].cpu().numpy()
        got_rot = sc.rotations[lo:lo + n].cpu().numpy()
        got_rest = sc.shs[lo:lo + n, 1:, :].cpu().numpy()
        got_dc = sc.shs[lo:lo + n, 0:1, :].cpu().numpy()
        np.testing.assert_allclose(got_xyz, want["xyz"], atol=2e-6)
        sign = np.sign((got_rot * want["rotation"]).sum(1, keepdims=True))
        np.testing.assert_allclose(got_rot * sign, want["rotation"], atol=2e-6)
        np.testing.assert_allclose(got_rest, want["features_rest"], atol=2e-6)
        np.testing.assert_array_equal(got_dc, objs[oid]["features_dc"])
        lo += n
    # environment rows untouched
    np.testing.assert_array_equal(sc.means3D[:sc.n_env].cpu().numpy(), env["xyz"])


def test_pose_incremental_deltas_equal_absolute_pose():
    """Dynamic mode (src/gs/pegasus_setup.py:178-193) applies per-frame deltas about the current
    centroid; the kernel applies the absolute pose about the canonical centroid. Same result."""
    from pegasus_b200 import ComposedScene, synth
    from pegasus_b200.sh_rotation import quat_xyzw_to_rotation
    env, objs = util.small_scene(n_env=500, n_obj=(900,), seed=41)
    traj = synth.drop_trajectory(1, 12, seed=3)
    tr = {"1": {str(f): {"t": list(traj[f, 0, :3]), "q": list(traj[f, 0, 3:])} for f in range(12)}}
    cur = objs[1]
    R0 = oracle.quat_xyzw_to_matrix(tr["1"]["0"]["q"]).astype(np.float32)
    cur = oracle.apply_transformation(cur, R0, np.asarray(tr["1"]["0"]["t"], np.float32))
    for ts in range(1, 12):
        Rd, td = oracle.dynamic_pose_delta(tr, 1, ts)
        cur = oracle.apply_transformation(cur, Rd, td)
    sc = ComposedScene(env, objs, oracle.generate_colors(1))
    sc.set_poses([(quat_xyzw_to_rotation(traj[11, 0, 3:]), traj[11, 0, :3])])
    torch.cuda.synchronize()
    n = 900
    np.testing.assert_allclose(sc.means3D[sc.n_env:].cpu().numpy(), cur["xyz"], atol=2e-5)
    got_rot = sc.rotations[sc.n_env:].cpu().numpy()
    sign = np.sign((got_rot * cur["rotation"]).sum(1, keepdims=True))
    np.testing.assert_allclose(got_rot * sign, cur["rotation"], atol=2e-5)
    np.testing.assert_allclose(sc.shs[sc.n_env:, 1:, :].cpu().numpy(), cur["features_rest"], atol=2e-5)


def _compare_frame(out, ref, colors, K_ids):
    rgb = out["color"].permute(1, 2, 0).cpu().numpy()
    depth = out["depth"].permute(1, 2, 0).cpu().numpy()
    assert np.abs(rgb - ref["rgb"]).max() <= RGB_TOL
    assert (np.abs(depth - ref["depth"]) / np.maximum(np.abs(ref["depth"]), 1e-6)).max() <= DEPTH_RTOL
    seg = out["seg_color"].permute(1, 2, 0).cpu().numpy()
    assert np.abs(seg - ref["seg_float"]).max() <= RGB_TOL
    # masks: exact outside the band where the reference's colour distance is within 1e-5 of 0.1
    vis = out["visible"].permute(1, 2, 0).cpu().numpy()
    sil = out["silhouette"].permute(1, 2, 0).cpu().numpy()
    sem = out["sem_seg"].cpu().numpy()
    stats = {}
    for ci, c in enumerate(colors):
        dist = np.linalg.norm(ref["seg_float"] - c, axis=2)
        band = np.abs(dist - 0.1) <= 1e-5
        bad = (vis[..., ci] != ref["visible"][..., ci]) & ~band
        assert not bad.any(), f"visible mask {ci}: {bad.sum()} mismatches outside the band"
    for oid in K_ids:
        ci = oid - 1
        bad = sil[..., ci] != ref["silhouette"][..., ci]
        stats[oid] = int(bad.sum())
    # sem-seg is uint8(255*seg): allow +-1 only where 255*seg is within 1e-3 of an integer
    diff = np.abs(sem.astype(np.int32) - ref["sem_seg"].astype(np.int32))
    frac = ref["seg_float"] * 255.0
    near = np.abs(frac - np.round(frac)) < 1e-3
    assert (diff[~near] == 0).all() and diff.max() <= 1
    return stats


@pytest.mark.parametrize("sizes,views", [((6000, 5000, 4000), 2), ((900,) * 18, 1)], ids=["3 objects", "18 objects"])
def test_composed_frame_matches_reference_k_plus_3_passes(sizes, views):
    """18 objects: past the object count up to which the compositing kernel runs its 5-CTAs-per-SM build."""
    from pegasus_b200 import ComposedScene, Camera, synth
    K = len(sizes)
    env, objs = util.small_scene(n_env=30000, n_obj=sizes, seed=51)
    colors = oracle.generate_colors(K)
    poses = _pose_list(K, 9)
    bg = np.zeros(3, np.float32)
    sc = ComposedScene(env, objs, colors, sh_mode="rotate")
    sc.set_poses(poses)
    # reference side: the oracle's K+3 passes on the object clouds the kernel just posed (isolates the
    # rasterizer from pose-kernel rounding, which has its own test)
    torch.cuda.synchronize()
    def rows(lo, hi):
        return dict(xyz=sc.means3D[lo:hi].cpu().numpy(), features_dc=sc.shs[lo:hi, 0:1, :].cpu().numpy(),
                    features_rest=sc.shs[lo:hi, 1:, :].cpu().numpy(), opacity=sc.opacity[lo:hi, None].cpu().numpy(),
                    scaling=sc.scales[lo:hi].cpu().numpy(), rotation=sc.rotations[lo:hi].cpu().numpy())
    env_act = rows(0, sc.n_env)
    posed = {}
    lo = sc.n_env
    for oid in sc.object_ids:
        n = objs[oid]["xyz"].shape[0]
        posed[oid] = rows(lo, lo + n)
        lo += n
    for c in synth.orbit_cameras(views, 640, 480, seed=3100):
        cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"])
        ocam = util.oracle_cam(c)
        # the product's Camera and the oracle's agree bit for bit on both matrices; the camera centre comes from a
        # 4x4 inverse (torch's LU vs numpy's: last ulp), so the oracle is given the product's
        assert np.array_equal(ocam["world_view_transform"], cam.world_view_transform.cpu().numpy())
        assert np.array_equal(ocam["full_proj_transform"], cam.full_proj_transform.cpu().numpy())
        ocam["camera_center"] = cam.camera_center.cpu().numpy()
        ref = oracle.render_frame_reference(ocam, env_act, posed, colors, bg, activated=True)
        out = sc.render(cam, torch.zeros(3, device="cuda"))
        np.testing.assert_array_equal(out["radii"].cpu().numpy(), ref["radii"])
        stats = _compare_frame(out, ref, colors, sc.object_ids)
        # main-pass images are expected bit-identical
        np.testing.assert_array_equal(out["color"].permute(1, 2, 0).cpu().numpy(), ref["rgb"])
        np.testing.assert_array_equal(out["seg_color"].permute(1, 2, 0).cpu().numpy(), ref["seg_float"])
        assert ref["visible"].sum() > 500 and ref["silhouette"].sum() > 500
        # silhouettes come from 1 - T_k instead of a 3-channel accumulation: only threshold-band pixels may differ
        assert sum(stats.values()) <= 7 * K, stats
        # the product's default numerics on the same frame: images inside the tolerances, masks equal to the exact
        # mode's except for a handful of threshold pixels
        fast = sc.render(cam, torch.zeros(3, device="cuda"), numerics="fast")
        np.testing.assert_array_equal(fast["radii"].cpu().numpy(), ref["radii"])
        assert_fast_close(fast["color"].permute(1, 2, 0).cpu().numpy(), ref["rgb"], "fast RGB", RGB_TOL)
        assert_fast_close(fast["depth"].permute(1, 2, 0).cpu().numpy(), ref["depth"], "fast depth", DEPTH_RTOL, rel=True)
        assert_fast_close(fast["seg_color"].permute(1, 2, 0).cpu().numpy(), ref["seg_float"], "fast seg", RGB_TOL)
        for name in ("visible", "silhouette"):
            assert int((fast[name] != out[name]).sum()) <= 4 * K, name
        assert int((fast["sem_seg"].int() - out["sem_seg"].int()).abs().max()) <= 1


def test_golden_mask_scene_from_reference_orchestration(golden):
    """The committed fixture: src/gs/render.py's own four helpers executed over the oracle rasterizer."""
    from pegasus_b200 import ComposedScene, Camera
    g = golden
    def cl(name):
        return dict(xyz=g[f"mask_{name}_xyz"], features_dc=g[f"mask_{name}_features_dc"],
                    features_rest=g[f"mask_{name}_features_rest"], opacity=g[f"mask_{name}_opacity"],
                    scaling=g[f"mask_{name}_scaling"], rotation=g[f"mask_{name}_rotation"])
    env = cl("env")
    objs = {int(k): cl(f"obj{int(k)}") for k in g["mask_obj_order"]}
    W, H = [int(v) for v in g["mask_WH"]]
    fovx, fovy = [float(v) for v in g["mask_cam_fov"]]
    sc = ComposedScene(env, objs, g["mask_colors"], sh_mode="canonical")   # identity pose
    cam = Camera(g["mask_cam_R"], g["mask_cam_T"], fovx, fovy, W, H)
    out = sc.render(cam, torch.zeros(3, device="cuda"))
    rgb = out["color"].permute(1, 2, 0).cpu().numpy()
    assert np.abs(rgb - g["mask_rgb"]).max() <= RGB_TOL
    depth = out["depth"].permute(1, 2, 0).cpu().numpy()
    assert (np.abs(depth - g["mask_depth"]) / np.maximum(np.abs(g["mask_depth"]), 1e-6)).max() <= DEPTH_RTOL
    vis = out["visible"].permute(1, 2, 0).cpu().numpy()
    sil = out["silhouette"].permute(1, 2, 0).cpu().numpy()
    assert (vis != g["mask_visible"]).sum() <= 2
    assert (sil != g["mask_silhouette"]).sum() <= 2
    assert np.abs(out["sem_seg"].cpu().numpy().astype(int) - g["mask_sem_seg"].astype(int)).max() <= 1


def test_render_helpers_return_reference_shapes():
    from pegasus_b200 import (ComposedScene, Camera, synth, render_rgb_and_depth, render_silhouette_mask,
                              render_visib_mask, render_semanticsegmentation_mask, render_frame)
    env, objs = util.small_scene(n_env=3000, n_obj=(800, 700), seed=61)
    colors = oracle.generate_colors(2)
    sc = ComposedScene(env, objs, colors)
    sc.set_poses(_pose_list(2, 3))
    c = synth.orbit_cameras(1, 200, 120, seed=5)[0]
    cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"])
    bg = torch.zeros(3, device="cuda")
    frame = render_frame(cam, sc, bg)
    rgb, depth = render_rgb_and_depth(cam, sc, None, bg, frame=frame)
    assert tuple(rgb.shape) == (120, 200, 3) and tuple(depth.shape) == (120, 200, 1) and not rgb.is_cuda
    sil = render_silhouette_mask(cam, sc, 200, 120, colors, None, bg, frame=frame)
    vis, seg = render_visib_mask(cam, sc, colors, 120, 200, None, bg, frame=frame)
    sem = render_semanticsegmentation_mask(cam, sc, colors, 120, 200, None, bg, False, frame=frame)
    assert sil.shape == (120, 200, 2) and sil.dtype == np.float64 and set(np.unique(sil)) <= {0.0, 1.0}
    assert vis.shape == (120, 200, 2) and tuple(seg.shape) == (120, 200, 3)
    assert sem.shape == (120, 200, 3) and sem.dtype == np.uint8
    # visible is a subset of silhouette (an occluded pixel is in the silhouette only), up to threshold pixels
    assert ((vis == 1) & (sil == 0)).sum() <= 5


@pytest.mark.parametrize("split", [False, True])
def test_pipelined_frames_on_three_streams_match_sequential(split):
    """Several frames in flight (one stream + workspace slot + output set each), a new pose every frame:
    the pose kernel of frame i+1 waits only for frame i's scene-read event.  With `split` every slot's
    compositing kernel runs on its own lower-priority stream (pg_launch_opts.composite_stream), forked from and
    joined back into the slot's stream.  Products must equal the one-frame-at-a-time results bit for bit."""
    from pegasus_b200 import ComposedScene, Camera, synth
    env, objs = util.small_scene(n_env=30000, n_obj=(5000, 4000), seed=71)
    colors = oracle.generate_colors(2)
    sc = ComposedScene(env, objs, colors)
    cams_h = synth.orbit_cameras(7, 640, 480, seed=17)
    cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"]) for c in cams_h]
    bg = torch.zeros(3, device="cuda")
    packets = [torch.from_numpy(sc.pack_poses(_pose_list(2, 100 + f)).numpy().copy()).cuda() for f in range(len(cams))]
    keys = ("color", "depth", "visible", "silhouette", "sem_seg", "radii")
    want = []
    for f, cam in enumerate(cams):
        sc.apply_pose_packets(packets[f])
        o = sc.render(cam, bg)
        want.append({k: o[k].clone() for k in keys})
    torch.cuda.synchronize()
    assert not torch.equal(want[0]["color"], want[1]["color"])
    n_slot = 3
    streams = [torch.cuda.Stream(priority=-1 if split else 0) for _ in range(n_slot)]
    comp_streams = [torch.cuda.Stream(priority=0) if split else None for _ in range(n_slot)]
    outs = [sc.alloc_outputs(640, 480) for _ in range(n_slot)]
    read_ev = [torch.cuda.Event() for _ in range(n_slot)]
    got = [None] * len(cams)
    for f, cam in enumerate(cams):
        sl = f % n_slot
        with torch.cuda.stream(streams[sl]):
            if got[f - n_slot] is None and f >= n_slot:
                pass
            if f >= n_slot:   # this slot's outputs are about to be overwritten: keep the previous frame's
                got[f - n_slot] = {k: outs[sl][k].clone() for k in keys}
            if f > 0:
                streams[sl].wait_event(read_ev[(f - 1) % n_slot])
            sc.apply_pose_packets(packets[f])
            sc.render(cam, bg, out=outs[sl], sync_check=False, slot=sl, scene_read_event=read_ev[sl],
                      composite_stream=comp_streams[sl])
    for f in range(len(cams) - n_slot, len(cams)):
        with torch.cuda.stream(streams[f % n_slot]):
            got[f] = {k: outs[f % n_slot][k].clone() for k in keys}
    torch.cuda.synchronize()
    for f in range(len(cams)):
        for k in keys:
            assert torch.equal(got[f][k], want[f][k]), (f, k)
    for sl in range(n_slot):
        assert sc.read_status(slot=sl)["overflow"] == 0


# ------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1] scale: ~3 M Gaussians, 1920x1080)
# ------------------------------------------------------------------------------------------------
def test_full_size_properties_1080p_3M():
    from pegasus_b200 import ComposedScene, Camera, synth
    from pegasus_b200.scene import export_binning
    env = synth.make_env(2_000_000, seed=1000)
    objs = {i + 1: synth.make_object(200_000, seed=2000 + i) for i in range(5)}
    colors = oracle.generate_colors(5)
    sc = ComposedScene(env, objs, colors)
    sc.set_poses(synth.static_poses(5, seed=4000))
    c = synth.orbit_cameras(4, 1920, 1080, seed=3000)[1]
    cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"])
    bg = torch.zeros(3, device="cuda")
    out = sc.render(cam, bg, reference_lists=True)
    R = out["num_rendered"]
    assert R > 1_000_000 and out["num_stored"] == R
    keys, plist, ranges = export_binning("cuda:0", sc.P, 1920, 1080, out["pair_capacity"], R)
    k = torch.from_numpy(keys.view(np.int64)).cuda()
    assert bool((k[1:] >= k[:-1]).all()), "sorted keys must be non-decreasing"
    # stable tie-break: equal keys keep ascending Gaussian index
    pl = torch.from_numpy(plist.view(np.int32)).cuda()
    eq = k[1:] == k[:-1]
    assert bool((pl[1:][eq] > pl[:-1][eq]).all())
    # ranges partition [0, R): non-empty tiles are consecutive runs
    rg = ranges.astype(np.int64)
    ne = rg[:, 1] > rg[:, 0]
    starts, ends = rg[ne, 0], rg[ne, 1]
    assert starts[0] == 0 and ends[-1] == R and (starts[1:] == ends[:-1]).all()
    tiles_of_keys = (keys >> np.uint64(32)).astype(np.int64)
    tile_ids = np.nonzero(ne)[0]
    assert (tiles_of_keys[starts] == tile_ids).all() and (tiles_of_keys[ends - 1] == tile_ids).all()
    # pair count == sum of tile rectangles of visible Gaussians, radii > 0 count == num_visible
    assert int((out["radii"] > 0).sum()) == out["num_visible"]
    # transmittance in (0,1], colour finite and >= 0, rendering is deterministic (idempotent)
    T = out["final_T"]
    assert float(T.min()) >= 0.0 and float(T.max()) <= 1.0
    assert bool(torch.isfinite(out["color"]).all()) and float(out["color"].min()) >= 0.0
    first = {kk: out[kk].clone() for kk in ("color", "depth", "visible", "silhouette", "sem_seg")}
    # the default path (only pairs that can contribute are stored) gives bit-identical products
    out2 = sc.render(cam, bg)
    assert out2["num_rendered"] == R and 0 < out2["num_stored"] < 0.8 * R
    for kk, v in first.items():
        assert torch.equal(v, out2[kk]), kk
    # masks=False path gives the same RGB/depth
    out3 = sc.render(cam, bg, masks=False)
    assert torch.equal(out3["color"], first["color"]) and torch.equal(out3["depth"], first["depth"])
    # a sample of tiles against the oracle compositor (cheap: oracle runs the full frame in seconds)
    sub = slice(0, sc.P)
    inp = dict(means3D=sc.means3D.cpu().numpy(), opacities=sc.opacity.cpu().numpy()[:, None],
               scales=sc.scales.cpu().numpy(), rotations=sc.rotations.cpu().numpy(), shs=sc.shs.cpu().numpy())
    ocam = util.oracle_cam(c)
    # the product's Camera and the oracle's agree bit for bit on both matrices; the camera centre comes from a 4x4
    # inverse (torch's LU vs numpy's: last ulp), so the oracle is given the product's
    assert np.array_equal(ocam["world_view_transform"], cam.world_view_transform.cpu().numpy())
    assert np.array_equal(ocam["full_proj_transform"], cam.full_proj_transform.cpu().numpy())
    ocam["camera_center"] = cam.camera_center.cpu().numpy()
    ref = util.oracle_forward(inp, ocam, np.zeros(3, np.float32))
    np.testing.assert_array_equal(out["radii"].cpu().numpy(), ref["radii"])
    assert R == ref["num_rendered"]
    np.testing.assert_array_equal(keys, ref["keys"])
    np.testing.assert_array_equal(plist, ref["point_list"])
    assert np.abs(out["color"].cpu().numpy() - ref["color"]).max() <= RGB_TOL
    assert (np.abs(out["depth"].cpu().numpy() - ref["depth"]) / np.maximum(np.abs(ref["depth"]), 1e-6)).max() <= DEPTH_RTOL
    # the FUSED mask passes at full size against the reference's K+3 passes (src/gs/render.py:36-129) on a band of
    # tile rows through the objects (the oracle clips every pass's rectangles to the band; inside it the lists, hence
    # the images and masks, are the complete ones)
    def rows(lo, hi):
        return dict(xyz=sc.means3D[lo:hi].cpu().numpy(), features_dc=sc.shs[lo:hi, 0:1, :].cpu().numpy(),
                    features_rest=sc.shs[lo:hi, 1:, :].cpu().numpy(), opacity=sc.opacity[lo:hi, None].cpu().numpy(),
                    scaling=sc.scales[lo:hi].cpu().numpy(), rotation=sc.rotations[lo:hi].cpu().numpy())
    posed, lo = {}, sc.n_env
    for oid in sc.object_ids:
        n = objs[oid]["xyz"].shape[0]
        posed[oid] = rows(lo, lo + n)
        lo += n
    sil_all = first["silhouette"].cpu().numpy()
    obj_rows = np.nonzero(sil_all.any(axis=(0, 2)))[0]
    assert obj_rows.size > 0
    mid = int(np.median(obj_rows)) // 16
    band = (max(mid - 2, 0), min(mid + 2, (1080 + 15) // 16))
    fr = oracle.render_frame_reference(ocam, rows(0, sc.n_env), posed, colors, np.zeros(3, np.float32), activated=True,
                                       band=band)
    y0, y1 = band[0] * 16, min(1080, band[1] * 16)
    crop = lambda a: a[y0:y1]
    np.testing.assert_array_equal(crop(first["color"].permute(1, 2, 0).cpu().numpy()), crop(fr["rgb"]))
    np.testing.assert_array_equal(crop(first["depth"].permute(1, 2, 0).cpu().numpy()), crop(fr["depth"]))
    vis = first["visible"].permute(1, 2, 0).cpu().numpy()
    sil = first["silhouette"].permute(1, 2, 0).cpu().numpy()
    assert crop(fr["silhouette"]).sum() > 1000 and crop(fr["visible"]).sum() > 1000
    for ci, cc in enumerate(colors):
        dist = np.linalg.norm(crop(fr["seg_float"]) - cc, axis=2)
        near = np.abs(dist - 0.1) <= 1e-5
        assert not ((crop(vis)[..., ci] != crop(fr["visible"])[..., ci]) & ~near).any(), ci
    # silhouettes come from 1 - T_k (closed form) instead of a 3-channel accumulation: threshold pixels only
    assert int((crop(sil) != crop(fr["silhouette"])).sum()) <= 40
    sem_d = np.abs(crop(first["sem_seg"].cpu().numpy()).astype(np.int32) - crop(fr["sem_seg"]).astype(np.int32))
    assert sem_d.max() <= 1 and int((sem_d > 0).sum()) <= 40

