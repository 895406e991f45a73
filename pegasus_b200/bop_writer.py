"""BOP-format dataset writer fed by packed frames.

Mirrors PegasusBOPDatasetWriter (/root/reference/src/tools/pegasus_working.py:298-592) and the
threaded `write_training_data` call of /root/reference/pegasus.py:333-365:

    <root>/<dataset>/camera.json
    <root>/<dataset>/train/<scene:06d>/{rgb,depth,mask,mask_visib,sem_mask}/<frame:06d>[_<idx:06d>].png
    <root>/<dataset>/train/<scene:06d>/scene_camera.json, scene_gt.json

Differences, all on the host and none in file content — with ONE exception: out-of-range values saturate
(`pg_pack_frame` clamps RGB * 255 to [0, 255] and depth * 1000 to [0, 65535]) where the reference's
`.astype("uint8")` / `.astype(np.uint16)` wrap modulo 256 / 65536 (pegasus.py:340-358), which turns a pixel brighter
than 1.0 dark; a deliberate deviation, pinned by tests/test_gpu_generate.py::test_pack_kernel_saturates.  Otherwise:
  * images arrive already packed by the GPU (`pg_pack_frame`: u8 RGB HWC, u16 depth in mm; masks
    u8 0/1 from the fused compositing pass), so no float image crosses PCIe and no numpy norm runs;
  * the per-object oriented bounding box is computed ONCE per object (ObjectMeta), not re-read from
    the OBJ mesh for every object of every frame (pegasus_working.py:469-471);
  * pose ground truth of all objects of a frame is one batched matrix product (`scene_gt_entries`).

Reference behaviour kept: `T_m2w` is built from R_init / t_init (pegasus_working.py:497-499), which
update_object_pose never refreshes — in dynamic mode the GT pose is the FIRST frame's.  Pass
``pose_source="current"`` to write the pose actually rendered instead.
"""
from __future__ import annotations

import json
import threading
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .cameras import focal2fov, fov2focal

# open3d box-point order -> the NDDS order the reference stores (pegasus_working.py:486-493)
_O3D_TO_NDDS = (0, 2, 5, 3, 1, 7, 4, 6)


@dataclass
class ObjectMeta:
    """What add_scene_gt_json needs to know about one object; box in MODEL coordinates."""
    obj_id: int                      # meta_info.ID (class id of the dataset)
    box_points: np.ndarray           # (8, 3) oriented-bounding-box corners, already in NDDS order
    box_center: np.ndarray           # (3,) centre of the box (projected as `projected_center`)
    mesh_center: np.ndarray = field(default_factory=lambda: np.zeros(3))  # `3d_bounding_center`

    @staticmethod
    def from_o3d_box(obj_id: int, o3d_box_points, box_center, mesh_center=None) -> "ObjectMeta":
        """Corners in open3d's get_box_points() order, re-ordered like the reference does."""
        p = np.asarray(o3d_box_points, dtype=np.float64)[list(_O3D_TO_NDDS)]
        c = np.asarray(box_center, dtype=np.float64)
        return ObjectMeta(obj_id, p, c, c if mesh_center is None else np.asarray(mesh_center, dtype=np.float64))

    @staticmethod
    def from_points(obj_id: int, xyz: np.ndarray) -> "ObjectMeta":
        """PCA-aligned box of a point set (e.g. the object cloud's means) when no mesh is at hand.
        open3d's minimal OBB of the mesh (the reference's source) is not reproduced here."""
        x = np.asarray(xyz, dtype=np.float64)
        c = x.mean(0)
        _, _, vt = np.linalg.svd(x - c, full_matrices=False)
        if np.linalg.det(vt) < 0:
            vt[2] = -vt[2]
        loc = (x - c) @ vt.T
        lo, hi = loc.min(0), loc.max(0)
        corners = np.array([[sx, sy, sz] for sx in (lo[0], hi[0]) for sy in (lo[1], hi[1]) for sz in (lo[2], hi[2])])
        pts = corners @ vt + c
        ctr = 0.5 * (lo + hi) @ vt + c
        return ObjectMeta(obj_id, pts, ctr, c)


def world_to_camera(cam_R: np.ndarray, cam_T: np.ndarray) -> np.ndarray:
    """T_w2c = [cam.R^T | cam.T] (pegasus_working.py:464-466); cam.R is stored transposed (COLMAP)."""
    T = np.eye(4)
    T[:3, :3] = np.asarray(cam_R, dtype=np.float64).T
    T[:3, 3] = np.asarray(cam_T, dtype=np.float64)
    return T


def scene_gt_entries(K: np.ndarray, cam_R, cam_T, bullet_ids: Sequence[int], metas: Sequence[ObjectMeta],
                     R_m2w: np.ndarray, t_m2w: np.ndarray) -> List[dict]:
    """scene_gt.json records of one frame (pegasus_working.py:457-566), all objects at once.
    R_m2w (n,3,3), t_m2w (n,3): model-to-world pose of each object.  Field for field:
    cam_R_m2c / cam_t_m2c = (T_w2c T_m2w)[:3], projections = K (T_w2c T_m2w)[:3] [p; 1] dehomogenised."""
    n = len(bullet_ids)
    T_w2c = world_to_camera(cam_R, cam_T)
    T_m2w = np.tile(np.eye(4), (n, 1, 1))
    T_m2w[:, :3, :3] = np.asarray(R_m2w, dtype=np.float64).reshape(n, 3, 3)
    T_m2w[:, :3, 3] = np.asarray(t_m2w, dtype=np.float64).reshape(n, 3)
    T = T_w2c[None] @ T_m2w                                   # (n,4,4)
    Pm = np.asarray(K, dtype=np.float64)[None] @ T[:, :3, :]  # (n,3,4)
    out = []
    for i in range(n):
        m = metas[i]
        pts = np.ones((9, 4))
        pts[:8, :3] = m.box_points
        pts[8, :3] = m.box_center
        hom = (Pm[i] @ pts.T).T                               # (9,3)
        # cv2.convertPointsFromHomogeneous: divide by w, except w == 0 -> scale 1
        w = np.where(hom[:, 2:3] != 0.0, hom[:, 2:3], 1.0)
        px = hom[:, :2] / w
        out.append({
            "cam_R_m2c": list(T[i, :3, :3].flatten()),
            "cam_t_m2c": list(T[i, :3, 3].flatten()),
            "T_w2c": list(T_w2c.flatten()),
            "T_m2w": list(T_m2w[i].flatten()),
            "obj_id": m.obj_id,
            "bullet_obj_id": bullet_ids[i],
            "3d_bounding_box_model_coord": np.asarray(m.box_points, dtype=np.float64).tolist(),
            "3d_bounding_center": np.asarray(m.mesh_center, dtype=np.float64).tolist(),
            "projected_center": px[8:9].tolist(),
            "projected_points": px[:8].tolist(),
        })
    return out


def _imwrite(path: str, img: np.ndarray) -> None:
    """PNG via OpenCV (imageio, the reference's writer, is not in this image).  RGB is swapped to
    OpenCV's BGR order so the file holds the same pixels imageio.imwrite would store."""
    import cv2
    if img.ndim == 3 and img.shape[2] == 3:
        img = img[:, :, ::-1]
    if not cv2.imwrite(path, np.ascontiguousarray(img)):
        raise IOError(f"could not write {path}")


class BOPDatasetWriter:
    """Same directory layout, file names and JSON fields as PegasusBOPDatasetWriter."""

    def __init__(self, dataset_name: str, dataset_output_path, fx: float, fy: float, image_width: int,
                 image_height: int, render_width: int, render_height: int, scene_id: int,
                 async_writes: bool = True):
        """fx, fy, image_width, image_height: the COLMAP intrinsics the reference reads from
        `camera_intr[1]` (pegasus_working.py:349-353)."""
        self.dataset_path = Path(dataset_output_path) / dataset_name
        self.dataset_path.mkdir(parents=True, exist_ok=True)
        self.render_width, self.render_height = int(render_width), int(render_height)
        self.model_path = self.dataset_path / "models"
        self.model_path.mkdir(parents=True, exist_ok=True)
        self._intr = (float(fx), float(fy), int(image_width), int(image_height))
        self.write_camera_json("camera.json")
        self.scene_id = scene_id
        self.scene_path = self.dataset_path / "train" / "{:06d}".format(scene_id)
        for name in ("depth", "mask_visib", "mask", "rgb", "sem_mask"):
            (self.scene_path / name).mkdir(parents=True, exist_ok=True)
        self.depth_path = self.scene_path / "depth"
        self.mask_visib_path = self.scene_path / "mask_visib"
        self.mask_path = self.scene_path / "mask"
        self.rgb_path = self.scene_path / "rgb"
        self.sem_mask_path = self.scene_path / "sem_mask"
        self.scene_camera_json_path = self.scene_path / "scene_camera.json"
        self.scene_gt_json_path = self.scene_path / "scene_gt.json"
        self.scene_camera_json: Dict = {}
        self.scene_gt_json: Dict[str, list] = {}
        self.async_writes = async_writes
        self._threads: List[threading.Thread] = []

    # ---- camera.json / scene_camera.json (pegasus_working.py:348-370, 440-455) ----
    def write_camera_json(self, file_name: str) -> None:
        fx0, fy0, w0, h0 = self._intr
        fx = fov2focal(focal2fov(fx0, w0), self.render_width)
        fy = fov2focal(focal2fov(fy0, h0), self.render_height)
        self.camera_json = {"cx": self.render_width / 2, "cy": self.render_height / 2, "depth_scale": 1.0,
                            "fx": fx, "fy": fy, "height": self.render_height, "width": self.render_width}
        with open(self.dataset_path / file_name, "w") as f:
            json.dump(self.camera_json, f, indent=4)

    def add_scene_camera_json(self, frame_id: int) -> None:
        K = np.eye(3, dtype=np.float64)
        K[0, 0], K[1, 1] = self.camera_json["fx"], self.camera_json["fy"]
        K[0, 2], K[1, 2] = self.camera_json["cx"], self.camera_json["cy"]
        self.scene_camera_json.update({frame_id: {"cam_K": list(K.flatten()), "depth_scale": 1.0}})
        self.K = K

    # ---- scene_gt.json (pegasus_working.py:457-566) ----
    def add_scene_gt_json(self, time_step: int, cam, bullet_ids: Sequence[int], metas: Sequence[ObjectMeta],
                          R_m2w: np.ndarray, t_m2w: np.ndarray) -> None:
        """cam: anything with .R (stored transposed) and .T.  R_m2w / t_m2w: see module docstring
        (R_init / t_init for reference behaviour, the current pose for pose_source="current")."""
        self.scene_gt_json.setdefault(str(time_step), [])
        self.scene_gt_json[str(time_step)].extend(
            scene_gt_entries(self.K, cam.R, cam.T, bullet_ids, metas, R_m2w, t_m2w))

    def write_scene_camera_json(self) -> None:
        with open(self.scene_camera_json_path, "w") as f:
            json.dump(self.scene_camera_json, f, indent=1)

    def write_scene_gt_json(self) -> None:
        with open(self.scene_gt_json_path, "w") as f:
            json.dump(self.scene_gt_json, f, indent=1)

    # ---- images (pegasus_working.py:407-438; pegasus.py:340-358) ----
    def write_training_data(self, frame_id: int, rgb_u8: Optional[np.ndarray] = None,
                            depth_u16: Optional[np.ndarray] = None, mask_visib: Optional[np.ndarray] = None,
                            mask_silhouette: Optional[np.ndarray] = None, sem_seg: Optional[np.ndarray] = None) -> None:
        """rgb_u8 (H,W,3) u8; depth_u16 (H,W) u16 millimetres; mask_* (n_colours,H,W) u8 in {0,1} — the
        planar layout the compositing kernel writes (the reference's (H,W,n) numpy masks, transposed);
        sem_seg (H,W,3) u8.  PNG encoding runs on a writer thread, as in pegasus.py:346."""
        args = (frame_id, rgb_u8, depth_u16, mask_visib, mask_silhouette, sem_seg)
        if not self.async_writes:
            return self._write(*args)
        th = threading.Thread(target=self._write, args=args)
        th.start()
        self._threads.append(th)

    def _write(self, frame_id, rgb_u8, depth_u16, mask_visib, mask_silhouette, sem_seg) -> None:
        name = "{:06d}.png".format(frame_id)
        if rgb_u8 is not None:
            _imwrite(str(self.rgb_path / name), rgb_u8)
        if sem_seg is not None:
            _imwrite(str(self.sem_mask_path / name), sem_seg)
        if depth_u16 is not None:
            _imwrite(str(self.depth_path / name), depth_u16)
        if mask_silhouette is not None:
            for idx in range(mask_silhouette.shape[0]):
                _imwrite(str(self.mask_path / "{:06d}_{:06d}.png".format(frame_id, idx)),
                         (mask_silhouette[idx] != 0).astype(np.uint8) * 255)
        if mask_visib is not None:
            for idx in range(mask_visib.shape[0]):
                _imwrite(str(self.mask_visib_path / "{:06d}_{:06d}.png".format(frame_id, idx)),
                         (mask_visib[idx] != 0).astype(np.uint8) * 255)

    def write_encoded(self, frame_id: int, streams: Dict, width: int, height: int, n_colours: int, rgb: bool = True,
                      sem_seg: bool = True, silhouette: bool = True, visible: bool = True) -> None:
        """The same files as `_write`, from zlib streams the GPU produced (pg_png_encode; `streams`: name -> bytes-like
        with the names of DatasetGenerator's encoder: rgb, depth, sem_seg, silhouette<k>, visible<k>).  The host only
        frames them: signature, IHDR, IDAT + CRC-32, IEND.  The files decode to the pixels `_write` stores."""
        from . import png_codec as pc
        name = "{:06d}.png".format(frame_id)
        if rgb:
            pc.write_png(str(self.rgb_path / name), pc.KIND_RGB8, width, height, streams["rgb"])
            pc.write_png(str(self.depth_path / name), pc.KIND_GRAY16, width, height, streams["depth"])
        if sem_seg:
            pc.write_png(str(self.sem_mask_path / name), pc.KIND_RGB8, width, height, streams["sem_seg"])
        for idx in range(n_colours):
            sub = "{:06d}_{:06d}.png".format(frame_id, idx)
            if silhouette:
                pc.write_png(str(self.mask_path / sub), pc.KIND_MASK8, width, height, streams[f"silhouette{idx}"])
            if visible:
                pc.write_png(str(self.mask_visib_path / sub), pc.KIND_MASK8, width, height, streams[f"visible{idx}"])

    def close(self) -> None:
        """Join the writer threads and flush both JSON files."""
        for th in self._threads:
            th.join()
        self._threads.clear()
        self.write_scene_camera_json()
        self.write_scene_gt_json()
