"""Drop-in replacement for the ``diff_gaussian_rasterization`` Python surface PEGASUS imports
(/root/reference/submodules/gaussian-splatting-pegasus/gaussian_renderer/__init__.py:14):
``GaussianRasterizationSettings`` (12-field NamedTuple, built by keyword at :38-51) and
``GaussianRasterizer`` whose forward is called by keyword at :87-95 and returns exactly
``(color[3,H,W], radii[P] int32, depth[1,H,W])``.

The work is done by libpegasus_b200.so through its C ABI (include/pegasus_b200.h); PyTorch only
owns the buffers and the stream.  No CPU path exists: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _Workspace:
    """Grow-only per-device scratch (the reference's geomBuffer/binningBuffer/imgBuffer role)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.capacity = 0  # pair capacity the layout was sized for
        self.key = None
        self.status_host = torch.zeros(_lib.STATUS_WORDS, dtype=torch.int32).pin_memory()

    def ensure(self, device, P, W, H, capacity):
        L = _lib.load()
        need = int(L.pg_workspace_bytes(P, W, H, capacity))
        if need == 0:
            raise RuntimeError("pg_workspace_bytes rejected the problem size")
        if self.buf is None or self.buf.device != device or self.buf.numel() < need:
            self.buf = None
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device)
                _lib.check(L.pg_workspace_init(C.c_void_p(self.buf.data_ptr()), self.buf.numel(),
                                               C.c_void_p(stream.cuda_stream)), "pg_workspace_init")
        self.capacity = capacity
        self.key = (P, W, H, capacity)
        return self.buf

    def status(self) -> dict:
        """The status words last copied into status_host (pg_read_status / pg_launch_opts.status_host)."""
        h = self.status_host
        return dict(num_rendered=int(h[0]) & 0xFFFFFFFF, overflow=int(h[1]), num_visible=int(h[2]),
                    num_stored=int(h[3]) & 0xFFFFFFFF, overflow_frames=int(h[4]) & 0xFFFFFFFF,
                    max_pairs_needed=int(h[5]) & 0xFFFFFFFF)


def grown_capacity(cap: int, st: dict) -> int:
    """Pair capacity after an overflow: the demand the library measured plus 1/8, at least double."""
    need = st["max_pairs_needed"]
    return int(min(max(need + need // 8 + 4096, min(2 * cap, 1 << 30), 1 << 20), 1 << 30))


_WORKSPACES = {}
_PAIR_CAPACITY_HINT = {}


def workspace_for(device, slot: int = 0) -> _Workspace:
    """Grow-only scratch of (device, slot).  Frames that are in flight at the same time (different CUDA
    streams) must use different slots; everything on one stream shares slot 0."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), int(slot))
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _WORKSPACES[key] = _Workspace()
    return ws


def default_pair_capacity(P: int, W: int, H: int) -> int:
    hint = _PAIR_CAPACITY_HINT.get((W, H), 0)
    cap = max(hint, 8 * P, 1 << 20)
    return int(min(cap, 1 << 30))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _prep(t: Optional[torch.Tensor], device, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: pegasus_b200 has no CPU path")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")
    return t.detach().to(torch.float32).contiguous()


def make_settings_struct(rs: GaussianRasterizationSettings, device, keep: list) -> _lib.RasterSettings:
    bg = _prep(rs.bg, device, "bg")
    view = _prep(rs.viewmatrix, device, "viewmatrix")
    proj = _prep(rs.projmatrix, device, "projmatrix")
    cam = _prep(rs.campos, device, "campos")
    keep.extend([bg, view, proj, cam])
    s = _lib.RasterSettings()
    s.image_height = int(rs.image_height)
    s.image_width = int(rs.image_width)
    s.tanfovx = float(rs.tanfovx)
    s.tanfovy = float(rs.tanfovy)
    s.bg = bg.data_ptr()
    s.scale_modifier = float(rs.scale_modifier)
    s.viewmatrix = view.data_ptr()
    s.projmatrix = proj.data_ptr()
    s.sh_degree = int(rs.sh_degree)
    s.campos = cam.data_ptr()
    s.prefiltered = int(bool(rs.prefiltered))
    s.debug = int(rs.debug)  # bit 0: sync after every stage, bit 1: collect compositing statistics
    return s


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings: GaussianRasterizationSettings, want_aux: bool = False,
                        sync_check: bool = True, reference_lists: bool = False, numerics=None):
    """Forward rasterization through the C ABI.  Returns (color, radii, depth, aux).

    reference_lists=True keeps the reference's complete (tile, Gaussian) pair lists (debug bit 2 of the
    C ABI) so that aux["n_contrib"] and pg_export_binning reproduce the reference's binning state; by
    default only the pairs that can contribute are stored and sorted (identical images)."""
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: pegasus_b200 has no CPU path")
    device = means3D.device
    L = _lib.load()
    keep = []
    with torch.cuda.device(device):
        s = make_settings_struct(raster_settings, device, keep)
        if reference_lists:
            s.debug |= 4
        P = int(means3D.shape[0])
        H, W = int(raster_settings.image_height), int(raster_settings.image_width)
        color = torch.zeros((3, H, W), dtype=torch.float32, device=device)
        depth = torch.zeros((1, H, W), dtype=torch.float32, device=device)
        radii = torch.zeros((P,), dtype=torch.int32, device=device)
        aux = {}
        if P == 0:
            return color, radii, depth, aux
        g = _lib.Gaussians()
        m3 = _prep(means3D, device, "means3D")
        sh_ = _prep(sh, device, "shs")
        cp_ = _prep(colors_precomp, device, "colors_precomp")
        op_ = _prep(opacities, device, "opacities")
        sc_ = _prep(scales, device, "scales")
        ro_ = _prep(rotations, device, "rotations")
        cv_ = _prep(cov3Ds_precomp, device, "cov3D_precomp")
        keep.extend([m3, sh_, cp_, op_, sc_, ro_, cv_])
        g.P = P
        g.means3D = m3.data_ptr()
        g.shs = sh_.data_ptr() if sh_ is not None and sh_.numel() else None
        g.sh_coeffs = int(sh_.shape[1]) if sh_ is not None and sh_.dim() == 3 else 0
        g.colors_precomp = cp_.data_ptr() if cp_ is not None and cp_.numel() else None
        g.opacities = op_.data_ptr()
        g.scales = sc_.data_ptr() if sc_ is not None and sc_.numel() else None
        g.rotations = ro_.data_ptr() if ro_ is not None and ro_.numel() else None
        g.cov3D_precomp = cv_.data_ptr() if cv_ is not None and cv_.numel() else None
        out = _lib.RasterOutputs()
        out.color, out.radii, out.depth = color.data_ptr(), radii.data_ptr(), depth.data_ptr()
        if want_aux:
            aux["final_T"] = torch.empty((H, W), dtype=torch.float32, device=device)
            out.final_T = aux["final_T"].data_ptr()
            if reference_lists:
                aux["n_contrib"] = torch.empty((H, W), dtype=torch.int32, device=device)
                out.n_contrib = aux["n_contrib"].data_ptr()
        ws = workspace_for(device)
        stream = torch.cuda.current_stream(device)
        cap = default_pair_capacity(P, W, H)
        opts = _lib.LaunchOpts()
        opts.numerics = _lib.numerics_code(numerics)
        if sync_check:
            opts.status_host = ws.status_host.data_ptr()
        while True:
            buf = ws.ensure(device, P, W, H, cap)
            rc = L.pg_rasterize_forward(C.byref(s), C.byref(g), C.byref(out), C.c_void_p(buf.data_ptr()),
                                        buf.numel(), cap, C.byref(opts), C.c_void_p(stream.cuda_stream))
            _lib.check(rc, "pg_rasterize_forward")
            if not sync_check:
                break
            stream.synchronize()
            st = ws.status()
            aux["num_rendered"], aux["num_visible"], aux["num_stored"] = st["num_rendered"], st["num_visible"], st["num_stored"]
            if not st["overflow"]:
                break
            if cap >= (1 << 30):
                raise RuntimeError("the (tile, Gaussian) pairs exceed the supported maximum of 2^30")
            cap = grown_capacity(cap, st)
            _PAIR_CAPACITY_HINT[(W, H)] = cap
        aux["pair_capacity"] = cap
        return color, radii, depth, aux


class GaussianRasterizer(nn.Module):
    sync_check = True  # read back R after each call and transparently grow the workspace on overflow
    reference_lists = False  # True: keep the reference's complete pair lists (aux["n_contrib"], export_binning)
    numerics = None  # None: the process default (pegasus_b200.set_numerics / PG_NUMERICS), else "exact" | "fast"

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self.aux = {}

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        if not positions.is_cuda:
            raise RuntimeError("positions must be a CUDA tensor: pegasus_b200 has no CPU path")
        with torch.no_grad(), torch.cuda.device(positions.device):
            L = _lib.load()
            pos = positions.detach().float().contiguous()
            view = self.raster_settings.viewmatrix.detach().float().contiguous()
            present = torch.zeros(pos.shape[0], dtype=torch.uint8, device=pos.device)
            stream = torch.cuda.current_stream(pos.device)
            _lib.check(L.pg_mark_visible(int(pos.shape[0]), C.c_void_p(pos.data_ptr()), C.c_void_p(view.data_ptr()),
                                         C.c_void_p(present.data_ptr()), C.c_void_p(stream.cuda_stream)),
                       "pg_mark_visible")
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        with torch.no_grad():
            color, radii, depth, aux = rasterize_gaussians(
                means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                self.raster_settings, want_aux=True, sync_check=self.sync_check,
                reference_lists=self.reference_lists, numerics=self.numerics)
        self.aux = aux
        return color, radii, depth
