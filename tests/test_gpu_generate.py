"""The dataset-generation loop (pegasus_b200/generate.py, mirror of pegasus.py:247-390) on the GPU:
pipelined frames must equal frames rendered one at a time (which the parity tests pin to the oracle),
packing must equal the reference's host conversions (golden from pegasus.py:345-355), the files on
disk must hold the same pixels, and sharding the frames over two ranks must produce the same dataset."""
import json
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _scene_and_path(n_frames=6, W=160, H=120):
    from pegasus_b200 import Camera, ComposedScene, synth
    from pegasus_b200.sh_rotation import quat_xyzw_to_rotation
    import oracle
    env, objs = util.small_scene(n_env=15000, n_obj=(2500, 2000))
    colors = oracle.generate_colors(2)
    scene = ComposedScene(env, objs, colors, device="cuda:0")
    cams_h = synth.orbit_cameras(n_frames, W, H, seed=3000)
    cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device="cuda:0") for c in cams_h]
    traj = synth.drop_trajectory(2, n_frames, seed=4000)
    poses = [[(quat_xyzw_to_rotation(traj[f, k, 3:]), traj[f, k, :3]) for k in range(2)] for f in range(n_frames)]
    return scene, cams, poses, objs


def _sequential(scene, cams, poses):
    """One frame at a time through the public API; host conversions exactly as pegasus.py:345-355."""
    bg = torch.zeros(3, device="cuda:0")
    out = []
    for f, cam in enumerate(cams):
        scene.set_poses(poses[f] if isinstance(poses[0], list) else poses)
        o = scene.render(cam, bg)
        rgb = o["color"].permute(1, 2, 0).cpu().numpy()
        depth = o["depth"].permute(1, 2, 0).cpu().numpy()
        # values above 1.0 (unclamped SH colour) wrap in numpy's uint8 cast; the packing kernel saturates instead
        out.append(dict(rgb=np.clip(np.ascontiguousarray(rgb) * np.float32(255), 0, 255).astype("uint8"),
                        depth=(depth * 1000).astype(np.uint16)[..., 0],
                        sem_seg=o["sem_seg"].cpu().numpy(), visible=o["visible"].cpu().numpy(),
                        silhouette=o["silhouette"].cpu().numpy()))
    return out


@pytest.mark.parametrize("W,H,n", [(1920, 8, 3), (333, 7, 2), (5, 3, 1), (16, 2, 4)])
def test_mask_bit_packing_round_trips_through_numpy_unpackbits(W, H, n):
    """pg_pack_masks: [n,H,W] u8 (0 / non-zero) -> [n,H,ceil(W/8)] bytes, pixel x in bit x % 8 — the wire format
    of the visible / silhouette masks; the writer thread expands it with numpy.unpackbits(bitorder="little")."""
    import ctypes as C
    from pegasus_b200 import _lib
    d = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(W * 131 + H)
    m = (torch.rand((n, H, W), generator=g) < 0.4).to(torch.uint8)
    m[0, 0, : min(W, 9)] = torch.tensor([1, 0, 0, 1, 1, 0, 1, 0, 1][: min(W, 9)], dtype=torch.uint8)
    m = m * torch.randint(1, 256, (n, H, W), generator=g).to(torch.uint8)  # any non-zero byte is "set"
    src = m.to(d)
    if W % 8:  # exercise the unaligned path as well
        src = torch.cat([torch.zeros(3, dtype=torch.uint8, device=d), src.reshape(-1)])[3:].reshape(n, H, W)
    Wb = (W + 7) // 8
    bits = torch.full((n, H, Wb), 0xAA, dtype=torch.uint8, device=d)
    L = _lib.load()
    st = torch.cuda.current_stream(d)
    _lib.check(L.pg_pack_masks(W, H, n, C.c_void_p(src.data_ptr()), C.c_void_p(bits.data_ptr()),
                               C.c_void_p(st.cuda_stream)), "pg_pack_masks")
    torch.cuda.synchronize()
    got = np.unpackbits(bits.cpu().numpy(), axis=-1, bitorder="little")[..., :W]
    np.testing.assert_array_equal(got, (m.numpy() != 0).astype(np.uint8))
    # padding bits of the last byte are zero
    if W % 8:
        assert int((bits.cpu().numpy()[..., -1] >> (W % 8)).max()) == 0


def test_pack_kernel_matches_reference_host_conversion():
    import ctypes as C
    from pegasus_b200 import _lib
    G = json.load(open(os.path.join(GOLD, "reference_host_rows.json")))
    rgb = np.array(G["pack_rgb_in"], dtype=np.float32)      # (H,W,3)
    depth = np.array(G["pack_depth_in"], dtype=np.float32)  # (H,W,1)
    H, W = rgb.shape[:2]
    d = torch.device("cuda", 0)
    color = torch.from_numpy(np.ascontiguousarray(rgb.transpose(2, 0, 1))).to(d)
    dep = torch.from_numpy(np.ascontiguousarray(depth.transpose(2, 0, 1))).to(d)
    o8 = torch.empty((H, W, 3), dtype=torch.uint8, device=d)
    o16 = torch.empty((H, W), dtype=torch.int16, device=d)
    L = _lib.load()
    st = torch.cuda.current_stream(d)
    _lib.check(L.pg_pack_frame(W, H, C.c_void_p(color.data_ptr()), C.c_void_p(dep.data_ptr()), C.c_void_p(o8.data_ptr()),
                               C.c_void_p(o16.data_ptr()), C.c_void_p(st.cuda_stream)), "pg_pack_frame")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o8.cpu().numpy(), np.array(G["pack_rgb_u8"], dtype=np.uint8))
    np.testing.assert_array_equal(o16.cpu().numpy().view(np.uint16), np.array(G["pack_depth_u16"], dtype=np.uint16))


def test_pack_kernel_saturates():
    """Out of range the reference's host casts wrap (uint8: modulo 256, uint16: modulo 65536); pg_pack_frame saturates
    instead — the one deliberate difference in file content (bop_writer.py header)."""
    import ctypes as C
    from pegasus_b200 import _lib
    d = torch.device("cuda", 0)
    H, W = 2, 4
    vals = torch.tensor([-0.2, 0.0, 0.5, 1.0, 1.004, 1.3, 2.0, 300.0], dtype=torch.float32)
    color = vals.reshape(1, H, W).repeat(3, 1, 1).contiguous().to(d)
    dep = torch.tensor([-1.0, 0.0, 0.0005, 1.2345, 65.535, 65.6, 100.0, 1e6], dtype=torch.float32).reshape(1, H, W).to(d)
    o8 = torch.empty((H, W, 3), dtype=torch.uint8, device=d)
    o16 = torch.empty((H, W), dtype=torch.int16, device=d)
    L = _lib.load()
    st = torch.cuda.current_stream(d)
    _lib.check(L.pg_pack_frame(W, H, C.c_void_p(color.data_ptr()), C.c_void_p(dep.data_ptr()), C.c_void_p(o8.data_ptr()),
                               C.c_void_p(o16.data_ptr()), C.c_void_p(st.cuda_stream)), "pg_pack_frame")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o8.cpu().numpy()[..., 0].reshape(-1), [0, 0, 127, 255, 255, 255, 255, 255])
    np.testing.assert_array_equal(o16.cpu().numpy().view(np.uint16).reshape(-1), [0, 0, 0, 1234, 65535, 65535, 65535, 65535])


@pytest.mark.parametrize("mode", ["dynamic", "static"])
def test_pipelined_generation_equals_sequential_frames(mode, tmp_path):
    import cv2
    from pegasus_b200 import BOPDatasetWriter, DatasetGenerator, ObjectMeta
    scene, cams, poses, objs = _scene_and_path()
    if mode == "static":
        poses = poses[-1]
    ref = _sequential(scene, cams, poses)
    W, H = cams[0].image_width, cams[0].image_height
    writer = BOPDatasetWriter("ds", tmp_path, 438.2178, 492.5640, 640, 480, W, H, scene_id=1, async_writes=False)
    metas = [ObjectMeta.from_points(10 + k, objs[oid]["xyz"]) for k, oid in enumerate(scene.object_ids)]
    seen = {}
    gen = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=2)
    stats = gen.generate(cams, poses=poses, writer=writer, metas=metas,
                         on_frame=lambda f, p: seen.__setitem__(f, {k: v.copy() for k, v in p.items()}))
    writer.close()
    assert stats["frames"] == len(cams) and sorted(seen) == list(range(len(cams)))
    root = tmp_path / "ds" / "train" / "000001"
    for f, r in enumerate(ref):
        for k in ("rgb", "depth", "sem_seg", "visible", "silhouette"):
            np.testing.assert_array_equal(seen[f][k], r[k], err_msg=f"{mode} frame {f} {k}")
        name = f"{f:06d}"
        assert np.array_equal(cv2.imread(str(root / "rgb" / f"{name}.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1], r["rgb"])
        assert np.array_equal(cv2.imread(str(root / "depth" / f"{name}.png"), cv2.IMREAD_UNCHANGED), r["depth"])
        for idx in range(2):
            assert np.array_equal(cv2.imread(str(root / "mask" / f"{name}_{idx:06d}.png"), cv2.IMREAD_UNCHANGED),
                                  r["silhouette"][idx] * 255)
            assert np.array_equal(cv2.imread(str(root / "mask_visib" / f"{name}_{idx:06d}.png"), cv2.IMREAD_UNCHANGED),
                                  r["visible"][idx] * 255)
    assert sum(r["visible"].sum() for r in ref) > 100  # the objects are in view
    gt = json.load(open(root / "scene_gt.json"))
    assert sorted(gt, key=int) == [str(f) for f in range(len(cams))]
    assert [e["obj_id"] for e in gt["0"]] == [10, 11] and [e["bullet_obj_id"] for e in gt["0"]] == scene.object_ids
    # reference behaviour: the pose written is R_init / t_init (first frame) for every frame
    first = poses[0] if mode == "dynamic" else poses
    for fkey in ("0", str(len(cams) - 1)):
        for k in range(2):
            Tm = np.array(gt[fkey][k]["T_m2w"]).reshape(4, 4)
            np.testing.assert_allclose(Tm[:3, :3], first[k][0], atol=1e-6)
            np.testing.assert_allclose(Tm[:3, 3], first[k][1], atol=1e-6)
    cam_json = json.load(open(root / "scene_camera.json"))
    assert len(cam_json) == len(cams) and cam_json["0"]["depth_scale"] == 1.0


def test_two_rank_sharding_writes_the_same_dataset(tmp_path):
    from pegasus_b200 import BOPDatasetWriter, DatasetGenerator
    from pegasus_b200.generate import merge_rank_fragments, write_rank_fragment
    scene, cams, poses, _ = _scene_and_path(n_frames=5)
    W, H = cams[0].image_width, cams[0].image_height
    single = {}
    DatasetGenerator(scene, W, H, frames_in_flight=2).generate(
        cams, poses=poses, on_frame=lambda f, p: single.__setitem__(f, p["rgb"].copy()))
    sharded = {}
    for rank in range(2):
        w = BOPDatasetWriter("ds", tmp_path, 438.2178, 492.5640, 640, 480, W, H, scene_id=0, async_writes=False)
        DatasetGenerator(scene, W, H, frames_in_flight=2).generate(
            cams, poses=poses, writer=w, rank=rank, world=2,
            on_frame=lambda f, p: sharded.__setitem__(f, p["rgb"].copy()))
        write_rank_fragment(w, rank)
    merge_rank_fragments(tmp_path / "ds" / "train" / "000000", world=2)
    assert sorted(sharded) == sorted(single) == list(range(5))
    for f in single:
        np.testing.assert_array_equal(sharded[f], single[f])
    cj = json.load(open(tmp_path / "ds" / "train" / "000000" / "scene_camera.json"))
    assert list(cj.keys()) == [str(f) for f in range(5)]
    assert not list((tmp_path / "ds" / "train" / "000000").glob("*.rank*.json"))


@pytest.mark.parametrize("mode", ["dynamic", "static"])
def test_gpu_png_dataset_decodes_to_the_host_encoded_dataset(mode, tmp_path):
    """png_on_gpu=True: the files come from zlib streams produced on the device (pg_png_encode) and framed on the
    host; every file decodes to the pixels of the dataset the host encoder writes, the JSON files are identical."""
    import cv2
    from pegasus_b200 import BOPDatasetWriter, DatasetGenerator, ObjectMeta
    scene, cams, poses, objs = _scene_and_path()
    if mode == "static":
        poses = poses[-1]
    W, H = cams[0].image_width, cams[0].image_height
    metas = [ObjectMeta.from_points(10 + k, objs[oid]["xyz"]) for k, oid in enumerate(scene.object_ids)]
    roots = {}
    for tag, on_gpu in (("host", False), ("gpu", True)):
        writer = BOPDatasetWriter(tag, tmp_path, 438.2178, 492.5640, 640, 480, W, H, scene_id=1, async_writes=False)
        gen = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=2, png_on_gpu=on_gpu)
        stats = gen.generate(cams, poses=poses, writer=writer, metas=metas)
        writer.close()
        assert stats["frames"] == len(cams)
        roots[tag] = tmp_path / tag / "train" / "000001"
        if on_gpu:
            assert gen.png_fallbacks == 0
            assert gen.d2h_bytes_per_frame < W * H * 8  # compressed streams, not raw products
    files = sorted(p.relative_to(roots["host"]) for p in roots["host"].rglob("*.png"))
    assert files == sorted(p.relative_to(roots["gpu"]) for p in roots["gpu"].rglob("*.png"))
    assert len(files) == len(cams) * (3 + 2 * 2)
    for rel in files:
        a = cv2.imread(str(roots["host"] / rel), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(str(roots["gpu"] / rel), cv2.IMREAD_UNCHANGED)
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), str(rel)
    for name in ("scene_gt.json", "scene_camera.json"):
        assert json.load(open(roots["host"] / name)) == json.load(open(roots["gpu"] / name))


def test_gpu_png_capacity_overflow_falls_back_per_frame(tmp_path):
    """Streams that outgrow the calibrated capacity are detected per frame and encoded again without a bound."""
    import cv2
    from pegasus_b200 import BOPDatasetWriter, DatasetGenerator
    scene, cams, poses, _ = _scene_and_path(n_frames=4)
    W, H = cams[0].image_width, cams[0].image_height
    ref = _sequential(scene, cams, poses)
    gen = DatasetGenerator(scene, W, H, frames_in_flight=2, writer_threads=2, png_on_gpu=True)
    gen.calibrate(cams, None)
    for e in gen.png_enc:  # far too small for the RGB stream
        caps = list(e.capacity)
        caps[0] = 256
        e.set_capacities(caps)
    sets = []
    while not gen._free.empty():
        sets.append(gen._free.get())
    for h in sets:
        h["png"] = torch.empty(gen.png_enc[0].arena_bytes, dtype=torch.uint8).pin_memory()
        gen._free.put(h)
    writer = BOPDatasetWriter("ds", tmp_path, 438.2178, 492.5640, 640, 480, W, H, scene_id=0, async_writes=False)
    gen.generate(cams, poses=poses, writer=writer)
    writer.close()
    assert gen.png_fallbacks == len(cams)
    root = tmp_path / "ds" / "train" / "000000"
    for f, r in enumerate(ref):
        assert np.array_equal(cv2.imread(str(root / "rgb" / f"{f:06d}.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1], r["rgb"])
        assert np.array_equal(cv2.imread(str(root / "depth" / f"{f:06d}.png"), cv2.IMREAD_UNCHANGED), r["depth"])


@pytest.mark.parametrize("png_on_gpu", [False, True])
def test_pair_capacity_overflow_regrows_and_resumes(png_on_gpu, tmp_path):
    """A pair capacity that is too small for some frames (calibration missed the worst view): the generator discards
    the frames in flight, grows the capacity to the measured demand, resumes at the failed frame — and the dataset is
    the one a sufficient capacity produces.  Nothing of an overflowed frame reaches the writer."""
    import cv2
    from pegasus_b200 import BOPDatasetWriter, DatasetGenerator
    scene, cams, poses, _ = _scene_and_path(n_frames=7)
    W, H = cams[0].image_width, cams[0].image_height
    ref = _sequential(scene, cams, poses)
    gen = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=2, png_on_gpu=png_on_gpu)
    full = gen.calibrate(cams, None)
    gen.pair_capacity = max(1 << 10, full // 3)  # too small for every frame; regrowth is clamped to >= 2^20 pairs
    gen._size_slots(cams[0])
    seen = []
    writer = BOPDatasetWriter("ds", tmp_path, 438.2178, 492.5640, 640, 480, W, H, scene_id=0, async_writes=False)
    stats = gen.generate(cams, poses=poses, writer=writer, on_frame=lambda f, p: seen.append(f))
    writer.close()
    assert stats["frames"] == len(cams) and stats["regrown"] >= 1 and gen.pair_capacity >= full // 3
    assert sorted(seen) == list(range(len(cams)))  # every frame delivered exactly once
    root = tmp_path / "ds" / "train" / "000000"
    for f, r in enumerate(ref):
        assert np.array_equal(cv2.imread(str(root / "rgb" / f"{f:06d}.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1], r["rgb"])
        assert np.array_equal(cv2.imread(str(root / "depth" / f"{f:06d}.png"), cv2.IMREAD_UNCHANGED), r["depth"])
        for idx in range(2):
            assert np.array_equal(cv2.imread(str(root / "mask_visib" / f"{f:06d}_{idx:06d}.png"), cv2.IMREAD_UNCHANGED),
                                  r["visible"][idx] * 255)
