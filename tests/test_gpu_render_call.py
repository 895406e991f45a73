"""The drop-in called the way the reference's own render() calls it.

tests/golden/render_calls.npz is a RECORDING (tools/make_golden_render_call.py): render()
(GSP/gaussian_renderer/__init__.py:19-103), the reference Camera (GSP/scene/cameras.py:17-57) and GaussianModel
accessors were executed from the reference's files with a recording stand-in for diff_gaussian_rasterization, in four
configurations (the call pegasus.py makes; convert_SHs_python; compute_cov3D_python with a scaling modifier;
override_color).  Here every recorded tensor is rebuilt on the GPU with the recorded shape, STRIDES, dtype and
requires_grad — the transposed, non-contiguous world_view_transform (strides (1, 4)), the camera centre as a strided
row of an inverse (stride 4), (P, 1) opacities, parameters that require grad — and handed to
diff_gaussian_rasterization keyword for keyword, in the recorded order.  The result must be the 3-tuple render()
unpacks and equal the CPU oracle on the same values (bit-identical with exact numerics)."""
import ast
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rebuild(g, key, dev):
    if key + "_none" in g.files:
        return None
    if key + "_scalar" in g.files:
        v = g[key + "_scalar"].item()
        return v
    shape, stride, dtype, req, is_param, contig = g[key + "_meta"]
    shape, stride = ast.literal_eval(str(shape)), ast.literal_eval(str(stride))
    vals = torch.from_numpy(g[key + "_values"])
    t = torch.empty_strided(shape, stride, dtype=getattr(torch, str(dtype)), device=dev)
    t.copy_(vals.to(dev))
    assert t.stride() == tuple(stride) and t.is_contiguous() == (str(contig) == "True")
    if str(is_param) == "True":
        t = torch.nn.Parameter(t, requires_grad=True)
    elif str(req) == "True":
        t.requires_grad_(True)
    return t


def _np(g, key):
    return None if key + "_none" in g.files else np.ascontiguousarray(g[key + "_values"])


@pytest.mark.parametrize("numerics", ["exact", "fast"])
def test_reference_render_calls_replayed_on_the_drop_in(numerics):
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from pegasus_b200 import set_numerics
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    g = np.load(os.path.join(ROOT, "tests", "golden", "render_calls.npz"))
    n_calls = int(g["n_calls"])
    assert n_calls == 4
    set_numerics(numerics)
    for i in range(n_calls):
        settings = {f: _rebuild(g, f"c{i}_settings_{f}", dev) for f in GaussianRasterizationSettings._fields}
        assert settings["viewmatrix"].stride() == (1, 4) and not settings["viewmatrix"].is_contiguous()
        assert settings["campos"].stride() == (4,)
        rs = GaussianRasterizationSettings(**settings)                      # keyword-built, as render() does (:38-51)
        rasterizer = GaussianRasterizer(raster_settings=rs)                 # :53
        order = [str(k) for k in g[f"c{i}_forward_order"]]
        assert order == ["means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp"]
        kwargs = {k: _rebuild(g, f"c{i}_forward_{k}", dev) for k in order}
        assert kwargs["opacities"].shape[1] == 1 and kwargs["means3D"].requires_grad
        rendered_image, radii, depth = rasterizer(**kwargs)                 # :87-95: exactly three tensors
        H, W = settings["image_height"], settings["image_width"]
        P = kwargs["means3D"].shape[0]
        assert rendered_image.shape == (3, H, W) and rendered_image.dtype == torch.float32 and rendered_image.is_cuda
        assert radii.shape == (P,) and radii.dtype == torch.int32
        assert depth.shape == (1, H, W) and depth.dtype == torch.float32
        assert (radii > 0).any()                                            # visibility_filter of :101

        # the oracle on the values the reference passed (viewmatrix / projmatrix as the transposed matrices they are)
        fw = lambda k: _np(g, f"c{i}_forward_{k}")
        st = lambda k: np.ascontiguousarray(g[f"c{i}_settings_{k}_values"])
        ref = oracle.rasterize_forward(fw("means3D"), fw("opacities"), st("viewmatrix"), st("projmatrix"), st("campos"),
                                       st("bg"), W, H, float(settings["tanfovx"]), float(settings["tanfovy"]),
                                       int(settings["sh_degree"]), shs=fw("shs"), colors_precomp=fw("colors_precomp"),
                                       scales=fw("scales"), rotations=fw("rotations"), cov3D_precomp=fw("cov3D_precomp"),
                                       scale_modifier=float(settings["scale_modifier"]))
        np.testing.assert_array_equal(radii.cpu().numpy(), ref["radii"])
        c, d = rendered_image.cpu().numpy(), depth.cpu().numpy()
        if numerics == "exact":
            np.testing.assert_array_equal(c, ref["color"])
            np.testing.assert_array_equal(d, ref["depth"])
        else:
            assert int((np.abs(c - ref["color"]) > 1e-3).sum()) <= 3
            rel = np.abs(d - ref["depth"]) / np.maximum(np.abs(ref["depth"]), 1e-6)
            assert int((rel > 1e-4).sum()) <= 3
        # the same call through contiguous, detached copies gives the same bits: the shim's .contiguous() is what
        # turns the reference's strided views into the flat [4 * col + row] layout the kernels read
        rs2 = GaussianRasterizationSettings(**{k: (v.detach().contiguous() if torch.is_tensor(v) else v)
                                               for k, v in settings.items()})
        c2, r2, d2 = GaussianRasterizer(raster_settings=rs2)(**{k: (None if v is None else v.detach().contiguous())
                                                               for k, v in kwargs.items()})
        assert torch.equal(c2, rendered_image) and torch.equal(r2, radii) and torch.equal(d2, depth)
