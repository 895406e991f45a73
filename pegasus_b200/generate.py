"""Dataset generation loop — the caller side of the hot path (mirrors
PEGASUS.generate_dataset, /root/reference/pegasus.py:247-390).

The reference, per frame: deep-copies the environment, merges every object into it, runs K+3
rasterizations, pulls K+2 float images to the host, thresholds colour distances with numpy, converts
to u8/u16 on the host, starts a PNG-writer thread and appends the pose ground truth.

Here, per frame: one H2D copy of the camera (+ the pose packets in dynamic mode), the pose kernel,
ONE fused frame (pg_render_composed), the packing kernels (pg_pack_frame: u8 RGB HWC, u16 depth mm;
pg_pack_masks: the 2 x n_colours mask planes as one bit per pixel), and D2H copies of the packed products
into a pinned host buffer set.  `frames_in_flight` frames
are pipelined, each on its own CUDA stream + workspace slot, so frame i+1's per-Gaussian and binning
stages overlap frame i's compositing and copies.  With `png_on_gpu=True` the frame's PNG streams are produced on
the device as well (pg_png_encode: Sub filter, run-length matches, Huffman coding with per-scene tables) and the
D2H copy carries the compressed streams instead of the raw products: the writer thread only frames them as PNG
files (CRC-32) — the host cost per frame drops from ~170 ms of libpng to a few ms, the D2H bytes to a third.  Host buffer sets are recycled through a free list:
a set goes back only after the writer (thread pool, like pegasus.py:346) is done with it, which
back-pressures rendering when PNG encoding is the bottleneck.

Multi-GPU (one process per GPU): frame f belongs to rank f % world; every rank writes its own image
files, and the JSON fragments are merged by rank 0 (`merge_rank_fragments`).
"""
from __future__ import annotations

import ctypes as C
import json
import queue
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import png_codec
from .png_gpu import FramePngEncoder, PngTables
from .rasterizer import grown_capacity
from .scene import ComposedScene
from .sh_rotation import POSE_WORDS, generate_pose_packets

DATA_POINTS = ("rgb", "seg_sil", "seg_vis", "sem_seg")  # pegasus.py: data_points of generate_dataset


class _CameraSlot:
    """Per-slot device staging of one camera (35 floats: view 16, full projection 16, centre 3)."""

    def __init__(self, device):
        self.buf = torch.zeros(35, dtype=torch.float32, device=device)
        self.world_view_transform = self.buf[0:16].view(4, 4)
        self.full_proj_transform = self.buf[16:32].view(4, 4)
        self.camera_center = self.buf[32:35]
        self.image_width = self.image_height = 0
        self.FoVx = self.FoVy = 0.0


def pack_cameras(cams: Sequence) -> torch.Tensor:
    """(n, 35) pinned float32: everything the renderer reads from a Camera, ready for async H2D."""
    host = torch.zeros((len(cams), 35), dtype=torch.float32)
    for j, cm in enumerate(cams):
        host[j, 0:16] = cm.world_view_transform.detach().cpu().contiguous().reshape(-1)
        host[j, 16:32] = cm.full_proj_transform.detach().cpu().contiguous().reshape(-1)
        host[j, 32:35] = cm.camera_center.detach().cpu()
    return host.pin_memory() if torch.cuda.is_available() else host


def _is_pose(p) -> bool:
    """True for one (R 3x3, t 3) pair, False for a list of such pairs."""
    try:
        return len(p) == 2 and np.shape(p[0]) == (3, 3) and np.shape(p[1]) == (3,)
    except (ValueError, TypeError):
        return False


class DatasetGenerator:
    """Renders a camera path over a ComposedScene and hands packed frame products to a writer."""

    def __init__(self, scene: ComposedScene, width: int, height: int, bg: Optional[torch.Tensor] = None,
                 frames_in_flight: int = 3, host_sets: Optional[int] = None, writer_threads: int = 4,
                 overlap_compositing: bool = False, numerics=None, png_on_gpu: bool = False):
        self.scene = scene
        self.dev = scene.device
        self.W, self.H = int(width), int(height)
        self.bg = bg if bg is not None else torch.zeros(3, device=self.dev)
        self.nslot = max(1, int(frames_in_flight))
        self.nc = int(scene.color_set.shape[0])
        self.K = len(scene.object_ids)
        self.writer_threads = max(1, int(writer_threads))
        self.numerics = numerics  # None: process default (pegasus_b200.set_numerics / PG_NUMERICS)
        dev, W, H, nc = self.dev, self.W, self.H, self.nc
        # overlap_compositing=True: each slot owns a HIGH-priority stream (pose, per-Gaussian stage, sorts, packing,
        # copies) and a normal-priority one for the compositing kernel (pg_launch_opts.composite_stream): the
        # latency-bound stages of frame i+1 take the SM slots frame i's compositing CTAs free.  Off by default: with
        # compositing at 5 CTAs per SM one stream per slot measures the same end to end and 3 % better device-resident.
        self.overlap = bool(overlap_compositing) and self.nslot > 1
        self.streams = [torch.cuda.Stream(device=dev, priority=-1 if self.overlap else 0) for _ in range(self.nslot)]
        self.comp_streams = [torch.cuda.Stream(device=dev, priority=0) if self.overlap else None
                             for _ in range(self.nslot)]
        self.outs = [scene.alloc_outputs(W, H, masks=True) for _ in range(self.nslot)]
        self.Wb = (W + 7) // 8  # bytes per row of a bit-packed mask plane (pg_pack_masks)
        self.packs = [dict(rgb=torch.empty((H, W, 3), dtype=torch.uint8, device=dev),
                           depth=torch.empty((H, W), dtype=torch.int16, device=dev))
                      for _ in range(self.nslot)]
        if not png_on_gpu:  # the bit-packed mask planes are the raw path's wire format
            for pk in self.packs:
                pk.update(visible=torch.empty((nc, H, self.Wb), dtype=torch.uint8, device=dev),
                          silhouette=torch.empty((nc, H, self.Wb), dtype=torch.uint8, device=dev))
        self.cam_dev = [_CameraSlot(dev) for _ in range(self.nslot)]
        self.pose_dev = [torch.zeros((max(self.K, 1), POSE_WORDS), dtype=torch.float32, device=dev)
                         for _ in range(self.nslot)]
        self.read_ev = [torch.cuda.Event() for _ in range(self.nslot)]
        self.done_ev = [torch.cuda.Event() for _ in range(self.nslot)]
        n_sets = host_sets if host_sets is not None else self.nslot + self.writer_threads
        self._free: "queue.Queue[Dict[str, torch.Tensor]]" = queue.Queue()
        for _ in range(max(n_sets, self.nslot)):
            # the frame's pg_status, copied by the library at the end of the frame
            hs = dict(status=torch.zeros(_lib.STATUS_WORDS, dtype=torch.int32).pin_memory())
            if not png_on_gpu:  # raw products; with PNG streams from the GPU the sets hold the streams (_calibrate_png)
                hs.update(rgb=torch.empty((H, W, 3), dtype=torch.uint8).pin_memory(),
                          depth=torch.empty((H, W), dtype=torch.int16).pin_memory(),
                          sem_seg=torch.empty((H, W, 3), dtype=torch.uint8).pin_memory(),
                          visible=torch.empty((nc, H, self.Wb), dtype=torch.uint8).pin_memory(),
                          silhouette=torch.empty((nc, H, self.Wb), dtype=torch.uint8).pin_memory())
            self._free.put(hs)
        self.h2d_bytes_per_frame = 35 * 4
        # u8 RGB + u16 depth + u8 sem-seg per pixel, the 2 x nc masks as ONE BIT per pixel: they are most of the
        # planes of a frame, and with 8 GPUs per host the D2H stream is what the host side saturates first
        self.d2h_bytes_per_frame = W * H * (3 + 2 + 3) + 2 * nc * H * self.Wb
        self.pair_capacity: Optional[int] = None
        self._overflow_seen: Dict[int, int] = {}  # slot -> sticky overflow_frames already reported
        # ---- PNG streams on the device (optional): per slot one encoder over the slot's product buffers
        self.png_on_gpu = bool(png_on_gpu)
        self.png_tables: Optional[PngTables] = None
        self.png_enc: List[FramePngEncoder] = []
        self.png_fallbacks = 0  # frames whose streams outgrew the calibrated capacity (encoded again, unbounded)
        if self.png_on_gpu:
            self.png_tables = PngTables(dev, ["rgb", "depth", "sem", "mask"])
            for sl in range(self.nslot):
                o, pk = self.outs[sl], self.packs[sl]
                imgs = [("rgb", png_codec.KIND_RGB8, "rgb", pk["rgb"]), ("depth", png_codec.KIND_GRAY16, "depth", pk["depth"]),
                        ("sem_seg", png_codec.KIND_RGB8, "sem", o["sem_seg"])]
                imgs += [(f"silhouette{k}", png_codec.KIND_MASK8, "mask", o["silhouette"][k]) for k in range(nc)]
                imgs += [(f"visible{k}", png_codec.KIND_MASK8, "mask", o["visible"][k]) for k in range(nc)]
                self.png_enc.append(FramePngEncoder(self.png_tables, W, H, imgs))

    # ------------------------------------------------------------------ capacity
    def calibrate(self, cams: Sequence, pose_packets: Optional[torch.Tensor] = None, margin: float = 1.05) -> int:
        """Untimed set-up: sizes every slot's workspace for the largest pair count over `cams`, so that the
        steady state never reads the count back.  Overflow is still detected (and reported) afterwards."""
        sc = self.scene
        max_R = 0
        for i, cam in enumerate(cams):
            if pose_packets is not None and self.K:
                sc.apply_pose_packets(pose_packets[i % pose_packets.shape[0]])
            o = sc.render(cam, self.bg, masks=True, out=self.outs[0], sync_check=True, numerics=self.numerics)
            max_R = max(max_R, o["num_stored"])
        self.pair_capacity = int(max_R * margin) + 4096
        self._size_slots(cams[0])
        if self.png_on_gpu:
            self._calibrate_png(cams, pose_packets)
        return self.pair_capacity

    def _size_slots(self, cam) -> None:
        """(Re)allocates every slot's workspace for the current pair capacity (one throw-away frame per slot)."""
        for sl in range(self.nslot):
            with torch.cuda.stream(self.streams[sl]):
                self.scene.render(cam, self.bg, masks=True, out=self.outs[sl], sync_check=True,
                                  pair_capacity=self.pair_capacity, slot=sl, numerics=self.numerics)
        torch.cuda.synchronize(self.dev)

    def _pack_slot(self, sl: int, st: torch.cuda.Stream) -> None:
        L, o = _lib.load(), self.outs[sl]
        _lib.check(L.pg_pack_frame(self.W, self.H, C.c_void_p(o["color"].data_ptr()), C.c_void_p(o["depth"].data_ptr()),
                                   C.c_void_p(self.packs[sl]["rgb"].data_ptr()), C.c_void_p(self.packs[sl]["depth"].data_ptr()),
                                   C.c_void_p(st.cuda_stream)), "pg_pack_frame")

    def _calibrate_png(self, cams: Sequence, pose_packets: Optional[torch.Tensor], margin: float = 1.3,
                       samples: int = 4) -> None:
        """Per-scene Huffman tables from the token histograms of a few sample views, then the stream capacities
        (= D2H bytes per frame) from the largest streams those views produce.  Statistics that drift later cost
        ratio, never correctness; a stream that outgrows its capacity is caught per frame and encoded again."""
        sc, enc, st = self.scene, self.png_enc[0], torch.cuda.current_stream(self.dev)
        pick = list(cams[::max(1, len(cams) // samples)])[:samples]

        def render(i, cam):
            if pose_packets is not None and self.K:
                sc.apply_pose_packets(pose_packets[i % pose_packets.shape[0]])
            sc.render(cam, self.bg, masks=True, out=self.outs[0], sync_check=True, pair_capacity=self.pair_capacity,
                      numerics=self.numerics)
            self._pack_slot(0, st)
        for i, cam in enumerate(pick):
            render(i, cam)
            enc.accumulate_hist(st)
        self.png_tables.rebuild_from_hist()
        sizes = np.zeros(len(enc.images), dtype=np.int64)
        for i, cam in enumerate(pick):
            render(i, cam)
            sizes = np.maximum(sizes, np.asarray(enc.measured_sizes(st)))
        # masks of one frame swap sizes with the view (coverage): all planes get the largest one's room
        names = enc.names
        mask_max = max([int(s) for n, s in zip(names, sizes) if n.startswith(("silhouette", "visible"))] or [0])
        caps = [int(margin * (mask_max if n.startswith(("silhouette", "visible")) else s)) + 8192 for n, s in zip(names, sizes)]
        for e in self.png_enc:
            e.set_capacities(caps)
        sets = []
        while not self._free.empty():
            sets.append(self._free.get())
        for h in sets:
            h["png"] = torch.empty(self.png_enc[0].arena_bytes, dtype=torch.uint8).pin_memory()
            h["png_result"] = torch.zeros((len(names), 2), dtype=torch.int32).pin_memory()
            self._free.put(h)
        self.d2h_bytes_per_frame = self.png_enc[0].arena_bytes + 8 * len(names)

    # ------------------------------------------------------------------ one frame, asynchronous
    def _issue(self, i: int, sl: int, cam, cam_row: torch.Tensor, pose_row: Optional[torch.Tensor], host: Dict,
               first: bool) -> None:
        sc, st = self.scene, self.streams[sl]
        L = _lib.load()
        with torch.cuda.stream(st):
            cs = self.cam_dev[sl]
            cs.buf.copy_(cam_row, non_blocking=True)
            cs.image_width, cs.image_height, cs.FoVx, cs.FoVy = cam.image_width, cam.image_height, cam.FoVx, cam.FoVy
            dynamic = pose_row is not None and self.K > 0
            if dynamic:
                self.pose_dev[sl].copy_(pose_row, non_blocking=True)
                if not first:  # the pose kernel rewrites rows the previous frame's per-Gaussian stage reads
                    st.wait_event(self.read_ev[(i - 1) % self.nslot])
                sc.apply_pose_packets(self.pose_dev[sl])
            o = self.outs[sl]
            sc.render(cs, self.bg, masks=True, out=o, sync_check=False, pair_capacity=self.pair_capacity, slot=sl,
                      scene_read_event=self.read_ev[sl] if dynamic else None, composite_stream=self.comp_streams[sl],
                      numerics=self.numerics, status_host=host["status"])
            self._pack_slot(sl, st)
            if self.png_on_gpu:
                self.png_enc[sl].encode(st)
                host["png"].copy_(self.png_enc[sl].arena, non_blocking=True)
                host["png_result"].copy_(self.png_enc[sl].result, non_blocking=True)
                self.done_ev[sl].record(st)
                return
            for name in ("visible", "silhouette"):
                _lib.check(L.pg_pack_masks(self.W, self.H, self.nc, C.c_void_p(o[name].data_ptr()),
                                           C.c_void_p(self.packs[sl][name].data_ptr()), C.c_void_p(st.cuda_stream)),
                           "pg_pack_masks")
            host["rgb"].copy_(self.packs[sl]["rgb"], non_blocking=True)
            host["depth"].copy_(self.packs[sl]["depth"], non_blocking=True)
            host["sem_seg"].copy_(o["sem_seg"], non_blocking=True)
            host["visible"].copy_(self.packs[sl]["visible"], non_blocking=True)
            host["silhouette"].copy_(self.packs[sl]["silhouette"], non_blocking=True)
            self.done_ev[sl].record(st)

    # ------------------------------------------------------------------ the loop
    def generate(self, cams: Sequence, poses: Optional[Sequence] = None, writer=None,
                 data_points: Sequence[str] = DATA_POINTS, metas: Optional[Sequence] = None,
                 pose_source: str = "reference", rank: int = 0, world: int = 1,
                 on_frame: Optional[Callable[[int, Dict[str, np.ndarray]], None]] = None,
                 pose_packets: Optional[torch.Tensor] = None,
                 frames: Optional[Sequence[int]] = None) -> Dict[str, int]:
        """cams[f]: camera of frame f.  poses: None (keep the scene's current pose), one list [(R, t)] * K
        (static mode: PegasusSetup.static_object_pose) or a list over frames of such lists (dynamic mode:
        dynamic_object_pose + update_object_pose, in absolute form).  `pose_packets` may pass the same
        thing as an already packed (frames, K, 103) HOST tensor (e.g. received by broadcast).
        frames: explicit frame list instead of the rank / world round robin.
        writer: a BOPDatasetWriter or None.  on_frame(f, products) runs on a writer thread and sees numpy
        views of the pinned set (the masks expanded from their bit-packed wire format), valid until it
        returns.  Returns counters."""
        for d in data_points:
            if d not in DATA_POINTS:
                raise ValueError(f"unknown data point {d!r}")
        sc = self.scene
        n = len(cams)
        # frames of this call: round-robin over the ranks, or an explicit list (a WorkItem of pegasus_b200.sweep:
        # a contiguous range of one scene's views, frame ids staying those of the whole camera path)
        mine = [int(f) for f in frames] if frames is not None else list(range(rank, n, world))
        if any(f < 0 or f >= n for f in mine):
            raise ValueError("frames must index cams")
        per_frame = None
        if pose_packets is not None:
            per_frame = pose_packets if pose_packets.dim() == 3 else pose_packets[None]
        elif poses is not None and self.K:
            seq = [poses] if _is_pose(poses[0]) else list(poses)
            per_frame = torch.from_numpy(np.stack([generate_pose_packets(p, sc.pivots, rotate_sh=(sc.sh_mode == "rotate"))
                                                   for p in seq]))
        if per_frame is not None:
            per_frame = per_frame.pin_memory() if not per_frame.is_pinned() else per_frame
            if per_frame.shape[0] not in (1, n):
                raise ValueError("poses must be given once (static) or once per frame (dynamic)")
        static = per_frame is not None and per_frame.shape[0] == 1
        if static and self.K:
            sc.pose_dev.copy_(per_frame[0], non_blocking=True)
            sc.apply_pose_packets(sc.pose_dev)
            torch.cuda.current_stream(self.dev).synchronize()
        cam_host = pack_cameras([cams[f] for f in mine])
        if self.pair_capacity is None:
            step = max(1, len(mine) // 16)
            self.calibrate([cams[f] for f in mine[::step]],
                           None if (per_frame is None or static) else per_frame[mine[::step]].to(self.dev))
        if self.png_on_gpu and not self.png_tables.calibrated:  # pair capacity was set by hand
            step = max(1, len(mine) // 16)
            self._calibrate_png([cams[f] for f in mine[::step]],
                                None if (per_frame is None or static) else per_frame[mine[::step]].to(self.dev))
        pool = ThreadPoolExecutor(max_workers=self.writer_threads) if (writer is not None or on_frame) else None
        want_rgb, want_sil = "rgb" in data_points, "seg_sil" in data_points
        want_vis, want_sem = "seg_vis" in data_points, "sem_seg" in data_points
        R0 = t0 = None
        if writer is not None and metas is not None and self.K:
            # reference behaviour: T_m2w from R_init / t_init, which update_object_pose never refreshes
            first_pose = self._pose_Rt(per_frame[0]) if per_frame is not None else None
            R0, t0 = first_pose if first_pose is not None else (None, None)
        inflight: List[Optional[tuple]] = [None] * self.nslot
        futures = []
        stats = dict(frames=0, overflow=0, regrown=0)
        # sticky overflow counters of the slots' workspaces before this call (calibration may have overflowed on purpose)
        self._overflow_seen = {sl: sc.read_status(slot=sl)["overflow_frames"] for sl in range(self.nslot)}

        def retire(sl):
            """Hands the slot's finished frame to the writer.  Returns None, or (sequence index, pairs needed) of a
            frame that exceeded the pair capacity: its products are invalid and never reach the writer."""
            job = inflight[sl]
            if job is None:
                return None
            f, cam, host, pose_row, seq = job
            self.done_ev[sl].synchronize()
            inflight[sl] = None
            if int(host["status"][1]):
                need = int(host["status"][5]) & 0xFFFFFFFF
                self._free.put(host)
                return seq, need
            stats["frames"] += 1
            W_ = self.W
            png_streams = None
            if self.png_on_gpu:
                if int(host["png_result"][:, 1].max()) != 0:
                    # statistics drifted past the calibrated capacity: encode this frame again without a bound while
                    # the slot's products are still intact (rare; costs a synchronisation)
                    self.png_fallbacks += 1
                    png_streams = self.png_enc[sl].encode_unbounded(self.streams[sl])
                else:
                    png_streams = self.png_enc[sl].streams(host["png"], host["png_result"])

            def make_products():
                # runs on the writer thread: the masks cross PCIe bit-packed and are expanded to the u8 0/1
                # planes the writer / callback expects only here
                unpack = lambda b: np.unpackbits(b.numpy(), axis=-1, bitorder="little")[..., :W_]
                return dict(rgb=host["rgb"].numpy() if want_rgb else None,
                            depth=host["depth"].numpy().view(np.uint16) if want_rgb else None,
                            sem_seg=host["sem_seg"].numpy() if want_sem else None,
                            visible=unpack(host["visible"]) if want_vis else None,
                            silhouette=unpack(host["silhouette"]) if want_sil else None)
            if writer is not None:
                writer.add_scene_camera_json(frame_id=f)
                if metas is not None and self.K:
                    if pose_source == "current" and pose_row is not None:
                        Rm, tm = self._pose_Rt(pose_row)
                    else:
                        Rm, tm = R0, t0
                    if Rm is not None:
                        writer.add_scene_gt_json(f, cam, list(sc.object_ids), metas, Rm, tm)

            def work():
                try:
                    if png_streams is not None:
                        if writer is not None:
                            writer.write_encoded(f, png_streams, W_, self.H, self.nc, rgb=want_rgb, sem_seg=want_sem,
                                                 silhouette=want_sil, visible=want_vis)
                        if on_frame is not None:
                            on_frame(f, {"png": png_streams})
                        return
                    prods = make_products()
                    if writer is not None:
                        writer._write(f, prods["rgb"], prods["depth"], prods["visible"], prods["silhouette"], prods["sem_seg"])
                    if on_frame is not None:
                        on_frame(f, prods)
                finally:
                    self._free.put(host)

            if pool is not None:
                futures.append(pool.submit(work))
            else:
                self._free.put(host)

        def regrow(seq, need):
            """A frame outgrew the pair capacity (the calibration views were not the worst): every frame in flight was
            issued after it and is discarded, the capacity grows to the measured demand (at least doubles), the
            slots' workspaces are re-sized, and rendering resumes AT that frame.  Nothing invalid was written."""
            for s2 in range(self.nslot):
                job = inflight[s2]
                if job is not None:
                    self.done_ev[s2].synchronize()
                    self._free.put(job[2])
                    inflight[s2] = None
            if self.pair_capacity >= (1 << 30):
                raise RuntimeError(f"frame {mine[seq]}: the (tile, Gaussian) pairs exceed the supported maximum of 2^30")
            self.pair_capacity = grown_capacity(self.pair_capacity, {"max_pairs_needed": need})
            stats["regrown"] += 1
            self._size_slots(cams[mine[seq]])
            self._overflow_seen = {s2: sc.read_status(slot=s2)["overflow_frames"] for s2 in range(self.nslot)}
            return seq

        i, n_mine, issued_first = 0, len(mine), False
        while True:
            sl = i % self.nslot
            bad = retire(sl)
            if bad is not None:
                i = regrow(*bad)
                continue
            if i >= n_mine:
                if all(j is None for j in inflight):
                    break
                i += 1  # drain: visit the remaining slots in issue order
                continue
            f = mine[i]
            host = self._free.get()  # blocks while every set is with the writer (back-pressure)
            pose_row = None if (per_frame is None or static) else per_frame[f]
            self._issue(i, sl, cams[f], cam_host[i], pose_row, host, first=not issued_first)
            issued_first = True
            inflight[sl] = (f, cams[f], host, pose_row if pose_row is not None else (per_frame[0] if per_frame is not None else None), i)
            i += 1
        for fu in futures:
            fu.result()
        if pool is not None:
            pool.shutdown()
        # belt and braces: the sticky counter of every slot's workspace covers all its frames since the last look
        for sl in range(self.nslot):
            ov = sc.read_status(slot=sl)["overflow_frames"]
            if ov != self._overflow_seen.get(sl, 0):
                self._overflow_seen[sl] = ov
                stats["overflow"] += 1
        if stats["overflow"]:
            raise RuntimeError("a frame exceeded the pair capacity without its status block saying so")
        return stats

    def _pose_Rt(self, packet: torch.Tensor):
        """(R (K,3,3), t (K,3)) of a packed pose row — the trajectory's own (R_k, t_k), which is what the
        reference stores as R_init / t_init and writes as T_m2w (pegasus_setup.py:165-171,
        pegasus_working.py:497-499)."""
        p = packet.detach().cpu().numpy().reshape(-1, POSE_WORDS).astype(np.float64)[: self.K]
        return p[:, 0:9].reshape(-1, 3, 3), p[:, 9:12]


def write_rank_fragment(writer, rank: int) -> None:
    """Every rank flushes its own JSON fragments next to the final files."""
    for name, data in (("scene_camera", writer.scene_camera_json), ("scene_gt", writer.scene_gt_json)):
        with open(Path(writer.scene_path) / f"{name}.rank{rank}.json", "w") as f:
            json.dump({str(k): v for k, v in data.items()}, f)


def merge_rank_fragments(scene_path, world: int, remove: bool = True) -> None:
    """Rank 0, after a barrier: merge the per-rank fragments into scene_camera.json / scene_gt.json with
    frames in ascending order (the order a single process would have written them in)."""
    scene_path = Path(scene_path)
    for name in ("scene_camera", "scene_gt"):
        merged = {}
        for r in range(world):
            p = scene_path / f"{name}.rank{r}.json"
            with open(p) as f:
                merged.update(json.load(f))
            if remove:
                p.unlink()
        ordered = {k: merged[k] for k in sorted(merged, key=lambda s: int(s))}
        with open(scene_path / f"{name}.json", "w") as f:
            json.dump(ordered, f, indent=1)
