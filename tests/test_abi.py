"""The C-ABI library loads without a GPU and exports every entry point include/pegasus_b200.h
declares; the ctypes mirror structs have the layout a C compiler gives the header's structs.
No compute call is made here (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from pegasus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pegasus_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/pegasus_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "pegasus_b200/_lib.py EXPORTS is out of sync with the header"
    assert b"sm_100a" in L.pg_version()


def test_ctypes_structs_match_the_header(tmp_path):
    structs = {"pg_raster_settings": _lib.RasterSettings, "pg_gaussians": _lib.Gaussians,
               "pg_raster_outputs": _lib.RasterOutputs, "pg_object_table": _lib.ObjectTable,
               "pg_frame_outputs": _lib.FrameOutputs, "pg_pose": _lib.Pose, "pg_canonical": _lib.Canonical,
               "pg_scene": _lib.Scene, "pg_status": _lib.Status, "pg_launch_opts": _lib.LaunchOpts,
               "pg_png_image": _lib.PngImage}
    prog = "#include <stdio.h>\n#include \"pegasus_b200.h\"\nint main(void){\n"
    for n in structs:
        prog += f'printf("{n} %zu\\n", sizeof({n}));\n'
    prog += 'printf("PG_NUM_STAGES %d\\nPG_MAX_OBJECTS %d\\nPG_MAX_COLORS %d\\n", PG_NUM_STAGES, PG_MAX_OBJECTS, PG_MAX_COLORS);return 0;}\n'
    c = tmp_path / "sz.c"
    c.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n, t in structs.items():
        assert int(out[n]) == C.sizeof(t), f"{n}: C {out[n]} bytes vs ctypes {C.sizeof(t)}"
    assert int(out["PG_NUM_STAGES"]) == _lib.NUM_STAGES == len(_lib.STAGE_NAMES)
    assert int(out["PG_MAX_OBJECTS"]) == _lib.PG_MAX_OBJECTS and int(out["PG_MAX_COLORS"]) == _lib.PG_MAX_COLORS
    assert C.sizeof(_lib.Pose) == 4 * _lib.POSE_WORDS == 412
    from pegasus_b200 import png_codec
    hdr = open(HEADER).read()
    assert f"#define PG_PNG_TABLE_WORDS {_lib.PNG_TABLE_WORDS}" in hdr and png_codec.TABLE_WORDS == _lib.PNG_TABLE_WORDS
    assert f"#define PG_PNG_HIST_WORDS {_lib.PNG_HIST_WORDS}" in hdr and png_codec.N_LITLEN == _lib.PNG_HIST_WORDS


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.load()
    assert L.pg_workspace_bytes(-1, 640, 480, 1 << 20) == 0
    assert L.pg_workspace_bytes(1000, 640, 480, 1 << 20) > 0
    rc = L.pg_rasterize_forward(None, None, None, None, 0, 0, None, None)
    assert rc == -1 and b"null" in L.pg_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "pg_rasterize_forward")
    assert L.pg_workspace_init(None, 0, None) == -3
    assert C.sizeof(_lib.Status) == 24 and _lib.STATUS_WORDS == 6


def test_launch_opts_are_validated_before_any_cuda_call():
    """pg_launch_opts is plain per-call data (no thread-local side channel): a composite stream needs both
    events, numerics must be a known mode.  Argument checks come first, so no GPU is needed."""
    L = _lib.load()
    s, g, out = _lib.RasterSettings(), _lib.Gaussians(), _lib.RasterOutputs()
    s.image_width, s.image_height, s.sh_degree = 64, 64, 3
    g.P, g.sh_coeffs = 0, 16
    g.shs, g.scales, g.rotations = 1, 1, 1  # non-null dummies: never dereferenced by the argument checks
    out.color = out.radii = out.depth = 1
    opts = _lib.LaunchOpts()
    opts.composite_stream = 1
    rc = L.pg_rasterize_forward(C.byref(s), C.byref(g), C.byref(out), C.c_void_p(1), 0, 1 << 20, C.byref(opts), None)
    assert rc == -1 and b"fork" in L.pg_last_error()
    opts = _lib.LaunchOpts()
    opts.numerics = 7
    rc = L.pg_rasterize_forward(C.byref(s), C.byref(g), C.byref(out), C.c_void_p(1), 0, 1 << 20, C.byref(opts), None)
    assert rc == -1 and b"numerics" in L.pg_last_error()
    opts.numerics = _lib.NUMERICS_FAST
    rc = L.pg_rasterize_forward(C.byref(s), C.byref(g), C.byref(out), C.c_void_p(1), 0, 1 << 20, C.byref(opts), None)
    assert rc == -3 and b"workspace too small" in L.pg_last_error()  # got past the option checks


def test_product_never_imports_the_oracle():
    for pkg in ("pegasus_b200", "diff_gaussian_rasterization"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f"{pkg}/{f} imports oracle"
                    assert "libpegasus_oracle" not in text
