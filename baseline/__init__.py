"""BASELINE (measurement + cross-check only; nothing under pegasus_b200/ imports this package).

`UpstreamStyleRasterizer` drives baseline/upstream_style.cu — a restatement of the public CUDA forward
rasterizer the reference pins as a submodule (absent from /root/reference), organised
the way upstream organises it: per-Gaussian preprocess, CUB inclusive scan, a blocking D2H read of
the pair count, 64-bit tile|depth keys, CUB radix sort, identifyTileRanges, 16x16-thread render.

`reference_frame` replays what the reference does with that rasterizer for ONE dataset frame
(/root/reference/pegasus.py:295-332 -> src/gs/render.py:14-129): K+3 forward passes — the merged
scene, every object alone, and the objects-only scene twice.  Only the rasterizer passes are run: the
reference's scene merges (deepcopy + vstack), torch activations and CPU-side numpy mask tests are left
out, which favours the baseline.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import build as _build

_LIB = None


def load():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path):
            _build.build()
        L = C.CDLL(path)
        L.base_create.restype = C.c_void_p
        L.base_destroy.argtypes = [C.c_void_p]
        L.base_forward.restype = C.c_longlong
        L.base_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
                                  [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 4 + [C.c_float, C.c_float] + \
                                  [C.c_void_p] * 4
        L.base_export.restype = C.c_int
        L.base_export.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


class UpstreamStyleRasterizer:
    def __init__(self):
        self.L = load()
        self.h = C.c_void_p(self.L.base_create())
        self.last_R = 0

    def __del__(self):
        try:
            self.L.base_destroy(self.h)
        except Exception:
            pass

    def forward(self, means3D, shs, opacities, scales, rotations, viewmatrix, projmatrix, campos, bg, W, H,
                tanfovx, tanfovy, sh_degree=3, scale_modifier=1.0, out=None):
        """All tensors float32 CUDA contiguous; shs (P, M, 3).  Returns dict(color, depth, radii, num_rendered)."""
        P = int(means3D.shape[0])
        dev = means3D.device
        if out is None:
            out = dict(color=torch.empty((3, H, W), dtype=torch.float32, device=dev),
                       depth=torch.empty((1, H, W), dtype=torch.float32, device=dev),
                       radii=torch.empty((max(P, 1),), dtype=torch.int32, device=dev))
        keep = [t.contiguous() for t in (viewmatrix, projmatrix, campos, bg)]
        stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            R = self.L.base_forward(self.h, P, int(sh_degree), int(shs.shape[1]), keep[3].data_ptr(), int(W), int(H),
                                    means3D.data_ptr(), shs.data_ptr(), opacities.data_ptr(), scales.data_ptr(),
                                    float(scale_modifier), rotations.data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(),
                                    keep[2].data_ptr(), float(tanfovx), float(tanfovy), out["color"].data_ptr(),
                                    out["depth"].data_ptr(), out["radii"].data_ptr(), C.c_void_p(stream.cuda_stream))
        if R < 0:
            raise RuntimeError("baseline rasterizer: CUDA error")
        self.last_R = int(R)
        out["num_rendered"] = int(R)
        return out

    def export(self, W, H):
        """Sorted 64-bit keys, point list and tile ranges of the last forward (device -> numpy)."""
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        dev = torch.device("cuda", torch.cuda.current_device())
        keys = torch.zeros(max(self.last_R, 1), dtype=torch.int64, device=dev)
        vals = torch.zeros(max(self.last_R, 1), dtype=torch.int32, device=dev)
        ranges = torch.zeros((tiles, 2), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev)
        rc = self.L.base_export(self.h, self.last_R, keys.data_ptr(), vals.data_ptr(), ranges.data_ptr(), tiles,
                                C.c_void_p(stream.cuda_stream))
        if rc:
            raise RuntimeError("baseline export failed")
        stream.synchronize()
        import numpy as np
        return (keys[:self.last_R].cpu().numpy().view(np.uint64), vals[:self.last_R].cpu().numpy().view(np.uint32),
                ranges.cpu().numpy().view(np.uint32))


def reference_frame(rast: UpstreamStyleRasterizer, scene, cam, bg, outs=None):
    """The reference's K+3 rasterizer passes for one frame on a pegasus_b200.ComposedScene's arrays
    (objects occupy the tail rows [n_env, P) in merge order).  Returns the list of pass outputs."""
    W, H = int(cam.image_width), int(cam.image_height)
    tx, ty = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
    first = [scene.n_env + int(v) for v in scene.first_rel]
    spans = [(0, scene.P)] + [(first[k], first[k + 1]) for k in range(len(first) - 1)] + [(scene.n_env, scene.P)] * 2
    res = []
    for i, (lo, hi) in enumerate(spans):
        o = rast.forward(scene.means3D[lo:hi], scene.shs[lo:hi], scene.opacity[lo:hi], scene.scales[lo:hi],
                         scene.rotations[lo:hi], cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                         bg, W, H, tx, ty, out=None if outs is None else outs[i])
        res.append(o)
    return res
