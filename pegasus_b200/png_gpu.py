"""Device-side bookkeeping of `pg_png_encode` (csrc/png.cu): which images a frame slot encodes, where their streams
go, the per-scene Huffman tables.  PyTorch only owns the buffers; the encoding is the library's, the table
construction and PNG framing are `png_codec`'s.

    tables = PngTables(device, ["rgb", "depth", "sem", "mask"])
    enc = FramePngEncoder(tables, W, H, [("rgb", KIND_RGB8, "rgb", rgb_u8), ("depth", KIND_GRAY16, "depth", d16), ...])
    enc.calibrate(stream)            # sample histogram of the current sources -> tables -> stream sizes -> capacities
    enc.encode(stream)               # every frame: 2 launches for all images
    host_arena.copy_(enc.arena); host_result.copy_(enc.result)      # one D2H each
    enc.streams(host_arena, host_result)   # name -> memoryview of the zlib stream (or raises PngOverflow)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import png_codec as pc


class PngOverflow(RuntimeError):
    """A stream did not fit the calibrated capacity (statistics drifted): encode that frame with
    `FramePngEncoder.encode_unbounded` or re-calibrate."""


class PngTables:
    """One static Huffman table per image group of a scene (device u32[G][PG_PNG_TABLE_WORDS]), shared by the slots."""

    def __init__(self, device, groups: Sequence[str]):
        self.device = device
        self.groups = list(groups)
        self.index = {g: i for i, g in enumerate(self.groups)}
        flat = np.ones(pc.N_LITLEN, np.int64)  # until calibrated: a flat code (valid, poor ratio)
        host = np.stack([pc.build_table(flat) for _ in self.groups])
        self.dev = torch.from_numpy(host.view(np.int32)).to(device)
        self.hist = torch.zeros((len(self.groups), _lib.PNG_HIST_WORDS), dtype=torch.int32, device=device)
        self.calibrated = False

    def table_ptr(self, group: str) -> int:
        return self.dev.data_ptr() + self.index[group] * _lib.PNG_TABLE_WORDS * 4

    def hist_ptr(self, group: str) -> int:
        return self.hist.data_ptr() + self.index[group] * _lib.PNG_HIST_WORDS * 4

    def rebuild_from_hist(self) -> None:
        """Histograms accumulated by calibration calls -> tables (synchronises)."""
        h = self.hist.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        host = np.stack([pc.build_table(h[i]) for i in range(len(self.groups))])
        self.dev.copy_(torch.from_numpy(host.view(np.int32)))
        self.hist.zero_()
        self.calibrated = True


class FramePngEncoder:
    """The images of ONE frame slot.  `images`: (name, kind, group, source tensor) with fixed device storage:
    u8 [H,W,3] (RGB8), int16/uint16 [H,W] (GRAY16), u8 [H,W] (MASK8; rows may be strided)."""

    def __init__(self, tables: PngTables, width: int, height: int, images: Sequence[Tuple[str, int, str, torch.Tensor]]):
        self.tables, self.W, self.H = tables, int(width), int(height)
        self.dev = tables.device
        self.images = list(images)
        self.names = [im[0] for im in self.images]
        n = len(self.images)
        L = _lib.load()
        self._L = L
        sb = int(L.pg_png_scratch_bytes(self.H))
        self.scratch = torch.empty((n, sb), dtype=torch.uint8, device=self.dev)
        self.result = torch.zeros((n, 2), dtype=torch.int32, device=self.dev)
        self.worst = [int(L.pg_png_worst_case_bytes(im[1], self.W, self.H)) for im in self.images]
        for (name, kind, group, t) in self.images:
            want = {pc.KIND_RGB8: (self.H, self.W, 3), pc.KIND_GRAY16: (self.H, self.W), pc.KIND_MASK8: (self.H, self.W)}[kind]
            if tuple(t.shape) != want or not t.is_cuda or t.stride(-1) != 1:
                raise ValueError(f"{name}: source must be a CUDA tensor of shape {want} with contiguous rows")
            if kind == pc.KIND_RGB8 and t.stride(1) != 3:
                raise ValueError(f"{name}: RGB pixels must be packed")
        self.capacity: List[int] = []
        self.offsets: List[int] = []
        self.arena = None
        self.set_capacities([64] * n)  # calibrate() sizes them

    # ------------------------------------------------------------------ buffers
    def set_capacities(self, caps: Sequence[int]) -> None:
        self.capacity = [min(max(64, (int(c) + 15) // 16 * 16), w) for c, w in zip(caps, self.worst)]
        self.offsets = list(np.concatenate([[0], np.cumsum(self.capacity)[:-1]]).astype(np.int64))
        self.arena_bytes = int(sum(self.capacity))
        self.arena = torch.empty(self.arena_bytes, dtype=torch.uint8, device=self.dev)
        self._descs = self._make_descs(self.arena, self.offsets, self.capacity, hist=False)
        self._descs_hist = None

    def _make_descs(self, arena: torch.Tensor, offsets, caps, hist: bool):
        arr = (_lib.PngImage * len(self.images))()
        for i, (name, kind, group, t) in enumerate(self.images):
            d = arr[i]
            d.src = t.data_ptr()
            d.kind = kind
            d.src_pitch = t.stride(0) * t.element_size()
            d.out = arena.data_ptr() + int(offsets[i])
            d.out_capacity = int(caps[i])
            d.table = self.tables.table_ptr(group)
            d.scratch = self.scratch.data_ptr() + i * self.scratch.shape[1]
            d.hist = self.tables.hist_ptr(group) if hist else None
            d.result = self.result.data_ptr() + 8 * i
        return arr

    # ------------------------------------------------------------------ per frame
    def encode(self, stream: torch.cuda.Stream) -> None:
        _lib.check(self._L.pg_png_encode(len(self.images), self._descs, self.W, self.H, C.c_void_p(stream.cuda_stream)),
                   "pg_png_encode")

    def streams(self, host_arena: torch.Tensor, host_result: torch.Tensor) -> Dict[str, memoryview]:
        """Host copies of arena / result -> {name: zlib stream}; raises PngOverflow when one did not fit."""
        res = host_result.numpy()
        if int(res[:, 1].max(initial=0)) != 0:
            raise PngOverflow("a PNG stream exceeded its calibrated capacity")
        buf = memoryview(host_arena.numpy())
        return {name: buf[int(o):int(o) + (int(res[i, 0]) & 0xFFFFFFFF)] for i, (name, o) in enumerate(zip(self.names, self.offsets))}

    # ------------------------------------------------------------------ set-up / rare paths
    def encode_unbounded(self, stream: torch.cuda.Stream) -> Dict[str, bytes]:
        """Synchronous encode of the current sources into worst-case sized buffers (cannot overflow)."""
        offs = list(np.concatenate([[0], np.cumsum(self.worst)[:-1]]).astype(np.int64))
        arena = torch.empty(int(sum(self.worst)), dtype=torch.uint8, device=self.dev)
        descs = self._make_descs(arena, offs, self.worst, hist=False)
        with torch.cuda.stream(stream):
            _lib.check(self._L.pg_png_encode(len(self.images), descs, self.W, self.H, C.c_void_p(stream.cuda_stream)),
                       "pg_png_encode")
            res = self.result.cpu().numpy()
            out = {}
            for i, name in enumerate(self.names):
                n = int(res[i, 0]) & 0xFFFFFFFF
                assert res[i, 1] == 0
                out[name] = arena[int(offs[i]):int(offs[i]) + n].cpu().numpy().tobytes()
        return out

    def accumulate_hist(self, stream: torch.cuda.Stream) -> None:
        """Adds the token histogram of the current sources to the tables' calibration histograms."""
        offs = list(np.concatenate([[0], np.cumsum(self.worst)[:-1]]).astype(np.int64))
        arena = torch.empty(int(sum(self.worst)), dtype=torch.uint8, device=self.dev)
        descs = self._make_descs(arena, offs, self.worst, hist=True)
        with torch.cuda.stream(stream):
            _lib.check(self._L.pg_png_encode(len(self.images), descs, self.W, self.H, C.c_void_p(stream.cuda_stream)),
                       "pg_png_encode")
        stream.synchronize()

    def measured_sizes(self, stream: torch.cuda.Stream) -> List[int]:
        """Stream sizes of the current sources with the current tables (synchronous, worst-case buffers)."""
        return [len(v) for v in self.encode_unbounded(stream).values()]
