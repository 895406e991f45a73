"""ctypes binding of libpegasus_b200.so (include/pegasus_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or a
CUDA call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

PG_MAX_OBJECTS = 32
PG_MAX_COLORS = 64

_f = C.c_void_p  # device pointers travel as integers


class RasterSettings(C.Structure):
    _fields_ = [("image_height", C.c_int32), ("image_width", C.c_int32), ("tanfovx", C.c_float),
                ("tanfovy", C.c_float), ("bg", _f), ("scale_modifier", C.c_float), ("viewmatrix", _f),
                ("projmatrix", _f), ("sh_degree", C.c_int32), ("campos", _f), ("prefiltered", C.c_int32),
                ("debug", C.c_int32)]


class Gaussians(C.Structure):
    _fields_ = [("P", C.c_int32), ("means3D", _f), ("shs", _f), ("sh_coeffs", C.c_int32),
                ("colors_precomp", _f), ("opacities", _f), ("scales", _f), ("rotations", _f),
                ("cov3D_precomp", _f)]


class RasterOutputs(C.Structure):
    _fields_ = [("color", _f), ("radii", _f), ("depth", _f), ("final_T", _f), ("n_contrib", _f)]


class ObjectTable(C.Structure):
    _fields_ = [("num_objects", C.c_int32), ("first", C.c_int32 * (PG_MAX_OBJECTS + 1)),
                ("color_index", C.c_int32 * PG_MAX_OBJECTS), ("num_colors", C.c_int32),
                ("colors", (C.c_float * 3) * PG_MAX_COLORS)]


class FrameOutputs(C.Structure):
    _fields_ = [("color", _f), ("radii", _f), ("depth", _f), ("final_T", _f), ("seg_color", _f),
                ("sem_seg", _f), ("visible", _f), ("silhouette", _f)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("pivot", C.c_float * 3), ("q", C.c_float * 4),
                ("D1", C.c_float * 9), ("D2", C.c_float * 25), ("D3", C.c_float * 49),
                ("rotate_sh", C.c_int32)]


POSE_WORDS = C.sizeof(Pose) // 4  # 103


class Canonical(C.Structure):
    _fields_ = [("n_total", C.c_int32), ("xyz", _f), ("rotation", _f), ("features_rest", _f)]


class Scene(C.Structure):
    _fields_ = [("P", C.c_int32), ("means3D", _f), ("rotations", _f), ("shs", _f)]


class Status(C.Structure):
    _fields_ = [("num_rendered", C.c_uint32), ("overflow", C.c_uint32), ("num_visible", C.c_uint32),
                ("num_stored", C.c_uint32), ("overflow_frames", C.c_uint32), ("max_pairs_needed", C.c_uint32)]


STATUS_WORDS = C.sizeof(Status) // 4  # 6

NUMERICS_EXACT, NUMERICS_FAST = 0, 1


class LaunchOpts(C.Structure):
    _fields_ = [("scene_read_event", _f), ("composite_stream", _f), ("fork_event", _f), ("join_event", _f),
                ("status_host", _f), ("numerics", C.c_int32)]


class PngImage(C.Structure):
    _fields_ = [("src", _f), ("kind", C.c_int32), ("src_pitch", C.c_int32), ("out", _f), ("out_capacity", C.c_uint32),
                ("table", _f), ("scratch", _f), ("hist", _f), ("result", _f)]


PNG_TABLE_WORDS = 610
PNG_HIST_WORDS = 286

EXPORTS = ["pg_version", "pg_last_error", "pg_workspace_bytes", "pg_workspace_init", "pg_rasterize_forward",
           "pg_render_composed", "pg_read_status", "pg_mark_visible", "pg_pose_apply",
           "pg_export_binning", "pg_pack_frame", "pg_profile_enable", "pg_profile_frames",
           "pg_profile_read", "pg_launch_count", "pg_read_stats", "pg_pack_masks", "pg_png_scratch_bytes",
           "pg_png_worst_case_bytes", "pg_png_encode"]

NUM_STAGES = 7
STAGE_NAMES = ["clear", "preprocess", "depth_sort", "emit", "tile_scan", "tile_sort", "composite"]

_LIB = None

# Process-wide default of pg_launch_opts.numerics for calls that do not name one (set_numerics() / PG_NUMERICS).
# "fast" by default: the reference's own CUDA expf is MUFU-based too, so neither mode is bit-identical to it, and both
# stay inside its tolerances; "exact" makes the images bit-reproducible on the CPU oracle (what the parity tests pin).
_NUMERICS = {"exact": NUMERICS_EXACT, "fast": NUMERICS_FAST}
_default_numerics = _NUMERICS.get(os.environ.get("PG_NUMERICS", "fast").lower(), NUMERICS_FAST)


def set_numerics(mode: str) -> None:
    """'exact': compositing bit-reproducible on the CPU oracle; 'fast': MUFU exp + one blend weight per Gaussian,
    within the reference tolerances (include/pegasus_b200.h, PG_NUMERICS_*)."""
    global _default_numerics
    _default_numerics = _NUMERICS[mode.lower()]


def numerics_code(mode=None) -> int:
    if mode is None:
        return _default_numerics
    if isinstance(mode, int):
        return int(mode)
    return _NUMERICS[mode.lower()]


def lib_path() -> str:
    return _build.OUT


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("PG_LIB_PATH")  # a tuning build of the same sources
    if not path:
        # (re)built when missing or when the sources it was built from have changed (content hash beside the .so);
        # raises if nvcc is unavailable or compilation fails — there is nothing to fall back to
        path = _build.build()
    L = C.CDLL(path)
    L.pg_version.restype = C.c_char_p
    L.pg_last_error.restype = C.c_char_p
    L.pg_workspace_bytes.restype = C.c_size_t
    L.pg_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
    L.pg_workspace_init.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.pg_workspace_init.restype = C.c_int
    L.pg_rasterize_forward.argtypes = [C.POINTER(RasterSettings), C.POINTER(Gaussians),
                                       C.POINTER(RasterOutputs), C.c_void_p, C.c_size_t, C.c_uint64,
                                       C.POINTER(LaunchOpts), C.c_void_p]
    L.pg_render_composed.argtypes = [C.POINTER(RasterSettings), C.POINTER(Gaussians), C.POINTER(ObjectTable),
                                     C.POINTER(FrameOutputs), C.c_void_p, C.c_size_t, C.c_uint64,
                                     C.POINTER(LaunchOpts), C.c_void_p]
    L.pg_read_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pg_mark_visible.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pg_pose_apply.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(Canonical), C.c_int32,
                                C.POINTER(Scene), C.c_void_p]
    L.pg_export_binning.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    L.pg_pack_frame.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]
    L.pg_pack_masks.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pg_pack_masks.restype = C.c_int
    L.pg_profile_enable.argtypes = [C.c_int32]
    L.pg_profile_frames.restype = C.c_int32
    L.pg_profile_read.argtypes = [C.c_int32, C.POINTER(C.c_float)]
    L.pg_launch_count.restype = C.c_uint64
    L.pg_read_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pg_png_scratch_bytes.argtypes = [C.c_int32]
    L.pg_png_scratch_bytes.restype = C.c_size_t
    L.pg_png_worst_case_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.pg_png_worst_case_bytes.restype = C.c_size_t
    L.pg_png_encode.argtypes = [C.c_int32, C.POINTER(PngImage), C.c_int32, C.c_int32, C.c_void_p]
    L.pg_png_encode.restype = C.c_int
    for name in ("pg_profile_enable", "pg_profile_read", "pg_read_stats"):
        getattr(L, name).restype = C.c_int
    for name in ("pg_rasterize_forward", "pg_render_composed", "pg_read_status", "pg_mark_visible",
                 "pg_pose_apply", "pg_export_binning", "pg_pack_frame"):
        getattr(L, name).restype = C.c_int
    _LIB = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
