"""Composed scene resident in HBM: one environment cloud + K object clouds in ONE pre-allocated
set of rasterizer-input arrays.

Replaces the per-frame ``copy.deepcopy(env)`` + six ``torch.vstack`` per object of
/root/reference/pegasus.py:255-264 and src/gs/gaussian_model.py:584-591: environment rows are
written once; every frame the pose kernel (pg_pose_apply) rewrites only the object rows' means,
quaternions and SH bands from the canonical (un-posed) clouds.  Row order is the reference's merge
order (environment first, then objects in dict order) because the index is the stable-sort tie-break.

Layout (all float32, contiguous):
    means3D (P,3)  shs (P,16,3)  opacity (P,)  scales (P,3)  rotations (P,4)
activations (sigmoid / exp / normalize of GSP/gaussian_renderer/__init__.py:57,67-68 via
src/gs/gaussian_model.py:105-125) are applied once at load with the same torch functions.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .rasterizer import (GaussianRasterizationSettings, default_pair_capacity, grown_capacity, make_settings_struct,
                         workspace_for, _PAIR_CAPACITY_HINT)
from .sh_rotation import POSE_WORDS, pose_packet

CLOUD_KEYS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")


def _t(a, device):
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


class ComposedScene:
    """env + objects -> rasterizer inputs.  `objects` is an ordered mapping bullet_id -> raw cloud
    dict (PLY-style, pre-activation parameters: keys CLOUD_KEYS)."""

    def __init__(self, env: Dict, objects: Dict[int, Dict], color_set, device="cuda", sh_mode: str = "rotate"):
        if not torch.cuda.is_available():
            raise RuntimeError("ComposedScene needs a CUDA device: pegasus_b200 has no CPU path")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if sh_mode not in ("rotate", "canonical"):
            raise ValueError("sh_mode must be 'rotate' or 'canonical'")
        if len(objects) > _lib.PG_MAX_OBJECTS:
            raise ValueError(f"at most {_lib.PG_MAX_OBJECTS} objects per scene")
        self.sh_mode = sh_mode
        self.object_ids: List[int] = list(objects.keys())
        dev = self.device
        n_env = int(np.asarray(env["xyz"]).shape[0]) if not isinstance(env["xyz"], torch.Tensor) else int(env["xyz"].shape[0])
        sizes = [int(objects[i]["xyz"].shape[0]) for i in self.object_ids]
        self.n_env = n_env
        self.first_rel = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)  # within the canonical arrays
        self.P = n_env + int(self.first_rel[-1])
        P = self.P
        self.means3D = torch.empty((P, 3), dtype=torch.float32, device=dev)
        self.shs = torch.empty((P, 16, 3), dtype=torch.float32, device=dev)
        self.opacity = torch.empty((P,), dtype=torch.float32, device=dev)
        self.scales = torch.empty((P, 3), dtype=torch.float32, device=dev)
        self.rotations = torch.empty((P, 4), dtype=torch.float32, device=dev)

        def put(lo, cloud, posed: bool):
            n = int(cloud["xyz"].shape[0])
            hi = lo + n
            self.shs[lo:hi, 0:1, :] = _t(cloud["features_dc"], dev).reshape(n, 1, 3)
            self.opacity[lo:hi] = torch.sigmoid(_t(cloud["opacity"], dev).reshape(n))
            self.scales[lo:hi] = torch.exp(_t(cloud["scaling"], dev).reshape(n, 3))
            if not posed:
                self.means3D[lo:hi] = _t(cloud["xyz"], dev)
                self.shs[lo:hi, 1:, :] = _t(cloud["features_rest"], dev).reshape(n, 15, 3)
                self.rotations[lo:hi] = torch.nn.functional.normalize(_t(cloud["rotation"], dev).reshape(n, 4))
            return hi

        put(0, env, posed=False)
        n_obj = int(self.first_rel[-1])
        self.canon_xyz = torch.empty((n_obj, 3), dtype=torch.float32, device=dev)
        self.canon_rot = torch.empty((n_obj, 4), dtype=torch.float32, device=dev)
        self.canon_rest = torch.empty((n_obj, 15, 3), dtype=torch.float32, device=dev)
        self.pivots = np.zeros((len(sizes), 3), dtype=np.float32)
        for k, oid in enumerate(self.object_ids):
            cl = objects[oid]
            lo, hi = int(self.first_rel[k]), int(self.first_rel[k + 1])
            put(n_env + lo, cl, posed=True)
            self.canon_xyz[lo:hi] = _t(cl["xyz"], dev)
            self.canon_rot[lo:hi] = _t(cl["rotation"], dev).reshape(-1, 4)
            self.canon_rest[lo:hi] = _t(cl["features_rest"], dev).reshape(-1, 15, 3)
            # centroid pivot of apply_rotation_on_xyz (src/gs/gaussian_model.py:488)
            self.pivots[k] = torch.mean(self.canon_xyz[lo:hi], 0).cpu().numpy()
        # colour set + table
        self.color_set = np.ascontiguousarray(
            color_set.detach().cpu().numpy() if isinstance(color_set, torch.Tensor) else color_set, dtype=np.float32)
        if self.color_set.shape[0] > _lib.PG_MAX_COLORS:
            raise ValueError(f"at most {_lib.PG_MAX_COLORS} colours")
        self.table = _lib.ObjectTable()
        self.table.num_objects = len(sizes)
        for k in range(len(sizes) + 1):
            self.table.first[k] = n_env + int(self.first_rel[k])
        for k, oid in enumerate(self.object_ids):
            ci = oid - 1
            if not (0 <= ci < self.color_set.shape[0]):
                raise ValueError(f"object id {oid} has no colour in a set of {self.color_set.shape[0]}")
            self.table.color_index[k] = ci
        self.table.num_colors = int(self.color_set.shape[0])
        for c in range(self.color_set.shape[0]):
            for ch in range(3):
                self.table.colors[c][ch] = float(self.color_set[c, ch])
        self._first_c = (C.c_int32 * (len(sizes) + 1))(*[int(v) for v in self.first_rel])
        self._pose_host = torch.zeros((max(len(sizes), 1), POSE_WORDS), dtype=torch.float32).pin_memory()
        self.pose_dev = torch.zeros((max(len(sizes), 1), POSE_WORDS), dtype=torch.float32, device=dev)
        self._status_pending = None
        # identity pose so that the object rows are valid before the first set_poses()
        if len(sizes):
            self.set_poses([(np.eye(3), np.zeros(3))] * len(sizes))

    # ---------------------------------------------------------------- poses
    def pack_poses(self, poses: Sequence) -> torch.Tensor:
        """poses: per object (R 3x3, t 3) ABSOLUTE pose w.r.t. the canonical cloud; returns the pinned
        host packet tensor (K, 103)."""
        assert len(poses) == len(self.object_ids)
        for k, (R, t) in enumerate(poses):
            self._pose_host[k] = torch.from_numpy(
                pose_packet(R, t, self.pivots[k], rotate_sh=(self.sh_mode == "rotate")))
        return self._pose_host

    def set_poses(self, poses: Sequence) -> None:
        self.pack_poses(poses)
        self.pose_dev.copy_(self._pose_host, non_blocking=True)
        self.apply_pose_packets(self.pose_dev)

    def apply_pose_packets(self, packets_dev: torch.Tensor) -> None:
        """packets_dev: (K,103) float32 DEVICE tensor — e.g. the buffer an NCCL broadcast filled."""
        K = len(self.object_ids)
        if K == 0:
            return
        assert packets_dev.is_cuda and packets_dev.dtype == torch.float32 and packets_dev.is_contiguous()
        L = _lib.load()
        canon = _lib.Canonical(int(self.first_rel[-1]), self.canon_xyz.data_ptr(), self.canon_rot.data_ptr(),
                               self.canon_rest.data_ptr())
        scene = _lib.Scene(self.P, self.means3D.data_ptr(), self.rotations.data_ptr(), self.shs.data_ptr())
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            _lib.check(L.pg_pose_apply(K, self._first_c, C.c_void_p(packets_dev.data_ptr()), C.byref(canon),
                                       self.n_env, C.byref(scene), C.c_void_p(stream.cuda_stream)), "pg_pose_apply")

    # ---------------------------------------------------------------- rendering
    def settings_for(self, cam, bg: torch.Tensor, sh_degree: int = 3, scaling_modifier: float = 1.0, debug=False):
        return GaussianRasterizationSettings(
            image_height=int(cam.image_height), image_width=int(cam.image_width),
            tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
            scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform,
            projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center,
            prefiltered=False, debug=debug)

    def alloc_outputs(self, W: int, H: int, masks: bool = True) -> Dict[str, torch.Tensor]:
        dev = self.device
        nc = int(self.color_set.shape[0])
        out = dict(color=torch.empty((3, H, W), dtype=torch.float32, device=dev),
                   depth=torch.empty((1, H, W), dtype=torch.float32, device=dev),
                   radii=torch.empty((self.P,), dtype=torch.int32, device=dev),
                   final_T=torch.empty((H, W), dtype=torch.float32, device=dev))
        if masks:
            out.update(seg_color=torch.empty((3, H, W), dtype=torch.float32, device=dev),
                       sem_seg=torch.empty((H, W, 3), dtype=torch.uint8, device=dev),
                       visible=torch.empty((nc, H, W), dtype=torch.uint8, device=dev),
                       silhouette=torch.empty((nc, H, W), dtype=torch.uint8, device=dev))
        return out

    def render(self, cam, bg: torch.Tensor, masks: bool = True, out: Optional[Dict] = None, sh_degree: int = 3,
               sync_check: bool = True, pair_capacity: Optional[int] = None, debug: int = 0,
               reference_lists: bool = False, slot: int = 0,
               scene_read_event: Optional[torch.cuda.Event] = None,
               composite_stream: Optional[torch.cuda.Stream] = None, numerics=None,
               status_host: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One frame: RGB + depth (+ seg render, sem-seg, visible and silhouette masks when
        masks=True) — everything the reference's K+3 passes produce (src/gs/render.py:14-129).

        The frame is enqueued on the current CUDA stream.  `slot` selects the workspace: frames in
        flight concurrently on different streams need distinct slots (and distinct `out` buffers).
        `scene_read_event` is recorded right after the per-Gaussian stage, the last reader of the scene
        arrays: the next frame's apply_pose_packets (on another stream) only has to wait for it.
        `composite_stream`: the compositing kernel runs there (pg_launch_opts.composite_stream), forked from and
        joined back into the current stream; give the current stream the higher priority and the
        following frame's per-Gaussian / sort stages co-run with this frame's compositing.
        `numerics`: "exact" | "fast" | None (process default).  `status_host`: pinned int32[6]; the frame's status
        block (overflow flag, pair counts) is copied there at the end of the frame, for callers that do not
        synchronise here (sync_check=False)."""
        L = _lib.load()
        H, W = int(cam.image_height), int(cam.image_width)
        if out is None:
            out = self.alloc_outputs(W, H, masks)
        keep = []
        with torch.cuda.device(self.device):
            s = make_settings_struct(self.settings_for(cam, bg, sh_degree, debug=debug), self.device, keep)
            if reference_lists:
                s.debug |= 4  # keep the reference's complete pair lists (export_binning parity)
            g = _lib.Gaussians(self.P, self.means3D.data_ptr(), self.shs.data_ptr(), 16, None,
                               self.opacity.data_ptr(), self.scales.data_ptr(), self.rotations.data_ptr(), None)
            fo = _lib.FrameOutputs(out["color"].data_ptr(), out["radii"].data_ptr(), out["depth"].data_ptr(),
                                   out["final_T"].data_ptr(),
                                   out["seg_color"].data_ptr() if masks else None,
                                   out["sem_seg"].data_ptr() if masks else None,
                                   out["visible"].data_ptr() if masks else None,
                                   out["silhouette"].data_ptr() if masks else None)
            ws = workspace_for(self.device, slot)
            stream = torch.cuda.current_stream(self.device)
            cap = pair_capacity or default_pair_capacity(self.P, W, H)
            table = self.table
            if not masks:
                table = _lib.ObjectTable()
                table.num_objects = 0
                table.num_colors = 0
            opts = _lib.LaunchOpts()
            opts.numerics = _lib.numerics_code(numerics)
            if scene_read_event is not None:
                scene_read_event.record(stream)  # creates the lazily-initialised handle; re-recorded by the library
                opts.scene_read_event = scene_read_event.cuda_event
            if composite_stream is not None:
                fork, join = _split_events(self.device, slot, stream)
                opts.composite_stream = composite_stream.cuda_stream
                opts.fork_event, opts.join_event = fork.cuda_event, join.cuda_event
            if sync_check:
                opts.status_host = ws.status_host.data_ptr()
            elif status_host is not None:
                opts.status_host = status_host.data_ptr()
            while True:
                buf = ws.ensure(self.device, self.P, W, H, cap)
                rc = L.pg_render_composed(C.byref(s), C.byref(g), C.byref(table), C.byref(fo),
                                          C.c_void_p(buf.data_ptr()), buf.numel(), cap, C.byref(opts),
                                          C.c_void_p(stream.cuda_stream))
                _lib.check(rc, "pg_render_composed")
                if not sync_check:
                    break
                stream.synchronize()
                st = ws.status()
                out["num_rendered"], out["num_visible"], out["num_stored"] = st["num_rendered"], st["num_visible"], st["num_stored"]
                if not st["overflow"]:
                    break
                if cap >= (1 << 30):
                    raise RuntimeError("the (tile, Gaussian) pairs exceed the supported maximum of 2^30")
                cap = grown_capacity(cap, st)
                _PAIR_CAPACITY_HINT[(W, H)] = cap
            out["pair_capacity"] = cap
        return out

    def read_stats(self, slot: int = 0) -> Dict[str, int]:
        """Compositing statistics of the last render(debug=2); synchronises."""
        L = _lib.load()
        ws = workspace_for(self.device, slot)
        host = torch.zeros(8, dtype=torch.int64).pin_memory()
        stream = torch.cuda.current_stream(self.device)
        _lib.check(L.pg_read_stats(C.c_void_p(ws.buf.data_ptr()), C.c_void_p(host.data_ptr()),
                                   C.c_void_p(stream.cuda_stream)), "pg_read_stats")
        stream.synchronize()
        return dict(pairs_evaluated=int(host[0]), pairs_exp=int(host[1]), pairs_blended=int(host[2]),
                    pixel_slots=int(host[3]), hits_env=int(host[4]), hits_obj_main=int(host[5]),
                    hits_obj_after=int(host[6]), cull_passes_after=int(host[7]))

    def read_status(self, slot: int = 0) -> Dict[str, int]:
        """Status of the last (possibly still running) render on this device, including the sticky fields
        (overflow_frames, max_pairs_needed) accumulated since the workspace was created; synchronises."""
        L = _lib.load()
        ws = workspace_for(self.device, slot)
        stream = torch.cuda.current_stream(self.device)
        _lib.check(L.pg_read_status(C.c_void_p(ws.buf.data_ptr()), C.c_void_p(ws.status_host.data_ptr()),
                                    C.c_void_p(stream.cuda_stream)), "pg_read_status")
        stream.synchronize()
        return ws.status()


_SPLIT_EVENTS: Dict = {}


def _split_events(device, slot: int, stream):
    """Fork / join events of one workspace slot (created once; torch creates the CUDA handle at the first
    record)."""
    key = (torch.device(device).index, slot)
    ev = _SPLIT_EVENTS.get(key)
    if ev is None:
        ev = (torch.cuda.Event(), torch.cuda.Event())
        for e in ev:
            e.record(stream)
        _SPLIT_EVENTS[key] = ev
    return ev


def export_binning(device, P: int, W: int, H: int, pair_capacity: int, num_rendered: int):
    """Tests/debug: sorted 64-bit keys, point list and tile ranges of the last forward; `num_rendered`
    is that forward's num_stored (== the reference's R when it ran with reference_lists=True)."""
    L = _lib.load()
    device = torch.device(device)
    ws = workspace_for(device)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    keys = torch.zeros(max(num_rendered, 1), dtype=torch.int64, device=device)
    plist = torch.zeros(max(num_rendered, 1), dtype=torch.int32, device=device)
    ranges = torch.zeros((tiles, 2), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        _lib.check(L.pg_export_binning(C.c_void_p(ws.buf.data_ptr()), P, W, H, pair_capacity,
                                       C.c_void_p(keys.data_ptr()), C.c_void_p(plist.data_ptr()),
                                       C.c_void_p(ranges.data_ptr()), C.c_void_p(stream.cuda_stream)),
                   "pg_export_binning")
        stream.synchronize()
    return (keys[:num_rendered].cpu().numpy().view(np.uint64), plist[:num_rendered].cpu().numpy().view(np.uint32),
            ranges.cpu().numpy().view(np.uint32))
