"""Host rows either side of the renderer against golden outputs of the
reference's OWN Python (tests/golden/reference_host_rows.json + ref_cloud.ply, generated in the build
container by tools/make_golden_host.py; /root/reference is never read here).  CPU only."""
import json
import os
import types

import numpy as np
import pytest

from pegasus_b200 import bop_writer, ply, trajectory
from pegasus_b200.bop_writer import BOPDatasetWriter, ObjectMeta

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def G():
    with open(os.path.join(GOLD, "reference_host_rows.json")) as f:
        return json.load(f)


def _extrinsics(G):
    return {k: types.SimpleNamespace(qvec=np.array(q), tvec=np.array(t))
            for k, q, t in zip(G["cam_ext_keys"], G["cam_ext_qvec"], G["cam_ext_tvec"])}


# ---- f3 / a-9: camera path (src/gs/pegasus_setup.py:85-143) ------------------------------------
@pytest.mark.parametrize("mode", ["random", "random+zoom", "sequence"])
def test_camera_path_matches_reference(G, mode):
    ref = G[f"campath_{mode}"]
    np.random.seed(1234)  # the reference draws from numpy's global generator
    poses = trajectory.camera_path_poses(_extrinsics(G), num_cameras=3, num_interpolation_steps=4, mode=mode)
    assert len(poses) == len(ref["R"]) == 12
    for (R, T), Rr, Tr in zip(poses, ref["R"], ref["T"]):
        # rotation goes matrix -> quaternion -> SLERP -> matrix in float64, stored float32
        np.testing.assert_allclose(R, np.array(Rr), atol=2e-7)
        np.testing.assert_allclose(T, np.array(Tr), atol=1e-7)


def test_camera_path_fov_quirks(G):
    """fx serves BOTH focal lengths and the FoVs come from the COLMAP image size (pegasus_setup.py:119-122)."""
    intr, ref = G["cam_intr"], G["campath_random"]
    np.random.seed(1234)
    cams = trajectory.create_camera_trajectory(_extrinsics(G), intr["fx"], intr["width"], intr["height"], 640, 480,
                                               num_cameras=3, num_interpolation_steps=4, device="cpu")
    assert len(cams) == 12
    assert cams[0].FoVx == pytest.approx(ref["FoVx"], abs=1e-15)
    assert cams[0].FoVy == pytest.approx(ref["FoVy"], abs=1e-15)
    assert [3, cams[0].image_height, cams[0].image_width] == ref["image_shape"]


def test_camera_path_rejects_unknown_mode(G):
    with pytest.raises(ValueError):
        trajectory.camera_path_poses(_extrinsics(G), mode="spiral")


# ---- f2: PLY layout (src/gs/gaussian_model.py:193-288) -----------------------------------------
KEYMAP = {"xyz": "_xyz", "features_dc": "_features_dc", "features_rest": "_features_rest", "opacity": "_opacity",
          "scaling": "_scaling", "rotation": "_rotation"}


def test_load_ply_reads_the_reference_file_like_the_reference(G):
    cloud = ply.load_ply(os.path.join(GOLD, "ref_cloud.ply"))
    assert ply.attribute_names(3) == G["ply_attribute_names"]
    for k, rk in KEYMAP.items():
        ref = np.array(G["ply" + rk], dtype=np.float32)
        assert cloud[k].dtype == np.float32 and cloud[k].shape == ref.shape, k
        assert np.array_equal(cloud[k], ref), k


def test_save_ply_writes_the_reference_bytes(G, tmp_path):
    """save_ply -> the same vertex block the reference's save_ply produced (header whitespace aside)."""
    src = os.path.join(GOLD, "ref_cloud.ply")
    cloud = ply.load_ply(src)
    out = str(tmp_path / "sub" / "cloud.ply")
    ply.save_ply(out, cloud)
    a, b = open(src, "rb").read(), open(out, "rb").read()
    assert a.split(b"end_header\n", 1)[1] == b.split(b"end_header\n", 1)[1]
    assert [l.split()[-1] for l in b.split(b"end_header\n")[0].splitlines() if l.startswith(b"property")] == \
           [n.encode() for n in G["ply_attribute_names"]]


def test_ply_errors_and_other_encodings(tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"plx\n")
    with pytest.raises(ValueError):
        ply.read_vertices(str(p))
    # truncated vertex block
    good = open(os.path.join(GOLD, "ref_cloud.ply"), "rb").read()
    p.write_bytes(good[:-8])
    with pytest.raises(ValueError):
        ply.read_vertices(str(p))
    # wrong SH degree for the file
    with pytest.raises(ValueError):
        ply.load_ply(os.path.join(GOLD, "ref_cloud.ply"), max_sh_degree=2)
    # ascii and big-endian encodings of the same table
    v = ply.read_vertices(os.path.join(GOLD, "ref_cloud.ply"))
    names = list(v.keys())
    table = np.stack([v[k] for k in names], axis=1)
    head = "ply\nformat {} 1.0\ncomment x\nelement vertex {}\n" + "".join(f"property float {k}\n" for k in names) + "end_header\n"
    pa = tmp_path / "a.ply"
    pa.write_bytes(head.format("ascii", len(table)).encode() +
                   "\n".join(" ".join(repr(float(x)) for x in row) for row in table).encode() + b"\n")
    pb = tmp_path / "b.ply"
    pb.write_bytes(head.format("binary_big_endian", len(table)).encode() + table.astype(">f4").tobytes())
    for q in (pa, pb):
        w = ply.read_vertices(str(q))
        for k in names:
            assert np.array_equal(np.asarray(w[k], dtype=np.float32), v[k]), (q.name, k)


def test_empty_cloud_round_trip(tmp_path):
    cloud = dict(xyz=np.zeros((0, 3), np.float32), features_dc=np.zeros((0, 1, 3), np.float32),
                 features_rest=np.zeros((0, 15, 3), np.float32), opacity=np.zeros((0, 1), np.float32),
                 scaling=np.zeros((0, 3), np.float32), rotation=np.zeros((0, 4), np.float32))
    path = str(tmp_path / "empty.ply")
    ply.save_ply(path, cloud)
    back = ply.load_ply(path)
    for k in cloud:
        assert back[k].shape == cloud[k].shape, k


# ---- f1 / f4 / a-11: BOP writer (src/tools/pegasus_working.py:298-592) --------------------------
def _writer(G, tmp_path, async_writes=False):
    intr = G["cam_intr"]
    return BOPDatasetWriter("ds", tmp_path, intr["fx"], intr["fy"], intr["width"], intr["height"], 640, 480,
                            scene_id=3, async_writes=async_writes)


def test_camera_json_and_scene_camera(G, tmp_path):
    w = _writer(G, tmp_path)
    assert open(tmp_path / "ds" / "camera.json").read() == G["bop_camera_json_text"]
    w.add_scene_camera_json(frame_id=4)
    assert {str(k): v for k, v in w.scene_camera_json.items()} == G["bop_scene_camera"]
    for d in ("rgb", "depth", "mask", "mask_visib", "sem_mask"):
        assert (tmp_path / "ds" / "train" / "000003" / d).is_dir()
    assert (tmp_path / "ds" / "models").is_dir()


def test_scene_gt_matches_reference_field_for_field(G, tmp_path):
    w = _writer(G, tmp_path)
    w.add_scene_camera_json(frame_id=4)
    objs = G["bop_objects"]
    metas = [ObjectMeta.from_o3d_box(o["obj_id"], o["o3d_box_points"], o["box_center"], o["mesh_center"]) for o in objs]
    cam = types.SimpleNamespace(R=np.array(G["bop_cam"]["R"]), T=np.array(G["bop_cam"]["T"]))
    w.add_scene_gt_json(4, cam, [o["bullet_id"] for o in objs], metas,
                        np.array([o["R_init"] for o in objs]), np.array([o["t_init"] for o in objs]))
    ref = G["bop_scene_gt"]
    assert list(w.scene_gt_json.keys()) == list(ref.keys()) == ["4"]
    assert len(w.scene_gt_json["4"]) == len(ref["4"]) == 3
    for got, want in zip(w.scene_gt_json["4"], ref["4"]):
        assert list(got.keys()) == list(want.keys())  # same fields in the same order
        for k in want:
            if isinstance(want[k], int):
                assert got[k] == want[k], k
            else:
                # one batched product vs. the reference's per-object chain: float64 round-off only
                np.testing.assert_allclose(np.array(got[k], dtype=np.float64), np.array(want[k], dtype=np.float64),
                                           rtol=1e-12, atol=1e-12, err_msg=k)
    # the flushed file parses to what was accumulated
    w.close()
    on_disk = json.load(open(tmp_path / "ds" / "train" / "000003" / "scene_gt.json"))
    assert on_disk["4"][0]["bullet_obj_id"] == ref["4"][0]["bullet_obj_id"]
    assert json.load(open(tmp_path / "ds" / "train" / "000003" / "scene_camera.json")) == G["bop_scene_camera"]


def test_projection_of_a_point_on_the_camera_plane_keeps_cv2_semantics():
    """cv2.convertPointsFromHomogeneous divides by w except when w == 0 (scale 1)."""
    K = np.array([[100.0, 0, 50], [0, 100, 40], [0, 0, 1]])
    meta = ObjectMeta(1, np.zeros((8, 3)), np.zeros(3))
    e = bop_writer.scene_gt_entries(K, np.eye(3), np.zeros(3), [1], [meta], np.eye(3)[None], np.zeros((1, 3)))[0]
    assert e["projected_center"] == [[0.0, 0.0]]


def test_images_round_trip_through_png(G, tmp_path):
    import cv2
    w = _writer(G, tmp_path, async_writes=True)
    rng = np.random.default_rng(5)
    H, W = 12, 20
    rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    sem = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    depth = rng.integers(0, 65536, size=(H, W), dtype=np.uint16)
    vis = rng.integers(0, 2, size=(2, H, W), dtype=np.uint8)
    sil = rng.integers(0, 2, size=(3, H, W), dtype=np.uint8)
    w.write_training_data(7, rgb_u8=rgb, depth_u16=depth, mask_visib=vis, mask_silhouette=sil, sem_seg=sem)
    w.close()
    root = tmp_path / "ds" / "train" / "000003"
    back = cv2.imread(str(root / "rgb" / "000007.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1]
    assert np.array_equal(back, rgb)  # stored RGB, as imageio would
    assert np.array_equal(cv2.imread(str(root / "sem_mask" / "000007.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1], sem)
    d = cv2.imread(str(root / "depth" / "000007.png"), cv2.IMREAD_UNCHANGED)
    assert d.dtype == np.uint16 and np.array_equal(d, depth)
    for i in range(3):
        m = cv2.imread(str(root / "mask" / f"000007_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(m, sil[i] * 255)
    for i in range(2):
        m = cv2.imread(str(root / "mask_visib" / f"000007_{i:06d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(m, vis[i] * 255)


def test_object_meta_from_points_is_a_tight_box():
    rng = np.random.default_rng(0)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    pts = (rng.uniform(-1, 1, size=(500, 3)) * [0.3, 0.1, 0.05]) @ R.T + [1, 2, 3]
    m = ObjectMeta.from_points(9, pts)
    assert m.box_points.shape == (8, 3) and m.obj_id == 9
    # every point lies inside the box: express in the box frame spanned by three edges from corner 0
    o = m.box_points[0]
    e = np.stack([m.box_points[4] - o, m.box_points[2] - o, m.box_points[1] - o])
    coords = (pts - o) @ np.linalg.pinv(e)
    assert coords.min() > -1e-9 and coords.max() < 1 + 1e-9
    np.testing.assert_allclose(m.box_center, m.box_points.mean(0), atol=1e-12)


# ---- multi-GPU: per-rank JSON fragments merged by rank 0 (pegasus_b200/generate.py) --------------
def test_rank_fragments_merge_in_frame_order(G, tmp_path):
    from pegasus_b200.generate import merge_rank_fragments, write_rank_fragment
    full = None
    for rank in range(3):
        w = _writer(G, tmp_path)
        for f in range(rank, 8, 3):
            w.add_scene_camera_json(frame_id=f)
            w.scene_gt_json[str(f)] = [{"obj_id": f}]
        write_rank_fragment(w, rank)
        full = w
    merge_rank_fragments(full.scene_path, world=3)
    cam = json.load(open(full.scene_path / "scene_camera.json"))
    gt = json.load(open(full.scene_path / "scene_gt.json"))
    assert list(cam.keys()) == list(gt.keys()) == [str(f) for f in range(8)]
    assert [gt[k][0]["obj_id"] for k in gt] == list(range(8))
    assert not list(full.scene_path.glob("*.rank*.json"))
