#!/usr/bin/env python
"""Record what the reference's own render() hands to the rasterizer.

Runs in the build container only (needs /root/reference).  render()
(GSP/gaussian_renderer/__init__.py:19-103), the Camera class (GSP/scene/cameras.py:17-57) and the GaussianModel
properties / activations (src/gs/gaussian_model.py:35-52, 105-128) are pulled out of the reference's files with `ast`
and executed as they are, with device="cuda" mapped to the CPU and a RECORDING stand-in for the
`diff_gaussian_rasterization` module: every GaussianRasterizationSettings field and every keyword of the
GaussianRasterizer call is written down with shape, stride, dtype, requires_grad and values.

tests/test_gpu_render_call.py rebuilds exactly those tensors on the GPU (same strides: the transposed, NON-contiguous
view matrix of cameras.py:54, the (P, 1) opacities, the fresh torch.cat of the SH features ...), calls
diff_gaussian_rasterization the same way, keyword for keyword, and compares with the CPU oracle.  The GPU box has no
/root/reference; the recording is committed as tests/golden/render_calls.npz.

    python tools/make_golden_render_call.py
"""
import ast
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (its import maps device="cuda" / .cuda() to the CPU and gives extract())

import torch  # noqa: E402
from torch import nn  # noqa: E402

REF, GSP = mg.REF, mg.GSP
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "render_calls.npz")

_orig_zeros_like = torch.zeros_like
torch.zeros_like = lambda *a, **k: _orig_zeros_like(*a, **mg._strip(k))


def extract_class(path, name, glb):
    tree = ast.parse(open(path).read())
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name][0]
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), glb)  # the class refers to itself by name
    return glb[name]


CALLS = []


def describe(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        return dict(shape=tuple(v.shape), stride=tuple(v.stride()), dtype=str(v.dtype).replace("torch.", ""),
                    requires_grad=bool(v.requires_grad), is_parameter=isinstance(v, nn.Parameter),
                    contiguous=bool(v.is_contiguous()), values=v.detach().contiguous().numpy().copy())
    return v


class RecSettings:
    FIELDS = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
              "sh_degree", "campos", "prefiltered", "debug")

    def __init__(self, **kw):
        assert tuple(kw.keys()) == self.FIELDS, kw.keys()   # the reference builds it by keyword, in this order
        self.kw = kw


class RecRasterizer:
    def __init__(self, raster_settings):
        self.s = raster_settings

    def __call__(self, **kw):
        CALLS.append(dict(settings={k: describe(v) for k, v in self.s.kw.items()},
                          forward={k: describe(v) for k, v in kw.items()}, forward_order=list(kw.keys())))
        P = kw["means3D"].shape[0]
        H, W = self.s.kw["image_height"], self.s.kw["image_width"]
        return torch.zeros(3, H, W), torch.zeros(P, dtype=torch.int32), torch.zeros(1, H, W)


def main():
    from utils import graphics_utils as gr, general_utils as gu, sh_utils
    render = mg.extract(os.path.join(GSP, "gaussian_renderer/__init__.py"), ["render"],
                        dict(torch=torch, math=math, GaussianRasterizationSettings=RecSettings,
                             GaussianRasterizer=RecRasterizer, GaussianModel=object, eval_sh=sh_utils.eval_sh))["render"]
    Camera = extract_class(os.path.join(GSP, "scene/cameras.py"), "Camera",
                           dict(torch=torch, nn=nn, np=np, getWorld2View2=gr.getWorld2View2,
                                getProjectionMatrix=gr.getProjectionMatrix))
    gm_path = os.path.join(REF, "src/gs/gaussian_model.py")
    glb = dict(torch=torch, nn=nn, np=np, build_scaling_rotation=gu.build_scaling_rotation,
               strip_symmetric=gu.strip_symmetric, inverse_sigmoid=gu.inverse_sigmoid)
    base = mg.extract(gm_path, ["setup_functions"], glb, cls="GaussianModelBase")
    props = mg.extract(gm_path, ["get_scaling", "get_rotation", "get_xyz", "get_features", "get_opacity", "get_covariance"],
                       glb, cls="GaussianModelBase")

    class PC:  # the attributes load_ply leaves behind (src/gs/gaussian_model.py:265-288), the reference's accessors
        pass
    PC.setup_functions = base["setup_functions"]
    for k, v in props.items():
        setattr(PC, k, v)

    rng = np.random.default_rng(77)
    P = 1200
    pc = PC()
    pc.setup_functions()
    pc.active_sh_degree = 3
    pc.max_sh_degree = 3
    par = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float).requires_grad_(True))
    xyz = rng.normal(size=(P, 3)) * np.array([0.5, 0.5, 0.15])
    pc._xyz = par(xyz)
    pc._features_dc = par(rng.uniform(-1.5, 1.5, size=(P, 1, 3)))
    pc._features_rest = par(rng.normal(size=(P, 15, 3)) * 0.1)
    pc._opacity = par(rng.normal(1.0, 2.0, size=(P, 1)))
    pc._scaling = par(np.log(rng.uniform(0.01, 0.06, size=(P, 3))))
    pc._rotation = par(rng.normal(size=(P, 4)))

    W, H = 200, 152
    eye = np.array([0.3, -1.6, 0.9])
    f = -eye / np.linalg.norm(eye)
    r_ = np.cross(f, [0, 0, 1.0]); r_ /= np.linalg.norm(r_)
    d_ = np.cross(f, r_)
    Rc = np.stack([r_, d_, f], axis=1)     # camera-to-world rotation = COLMAP R transposed (cameras.py stores this)
    Tc = -Rc.T @ eye
    fovx = math.radians(62.0)
    fovy = gr.focal2fov(gr.fov2focal(fovx, W), H)
    cam = Camera(colmap_id=0, R=Rc, T=Tc, FoVx=fovx, FoVy=fovy, image=torch.zeros(3, H, W), gt_alpha_mask=None,
                 image_name="golden", uid=0, data_device="cpu")

    class Pipe:
        debug = False
        convert_SHs_python = False
        compute_cov3D_python = False

    bg = torch.tensor([0.0, 0.0, 0.0], dtype=torch.float32)
    render(cam, pc, Pipe, bg)                                         # the call pegasus.py makes (src/gs/render.py:16)
    Pipe.convert_SHs_python = True
    render(cam, pc, Pipe, torch.tensor([1.0, 1.0, 1.0]))              # white background, colours precomputed in Python
    Pipe.convert_SHs_python = False
    Pipe.compute_cov3D_python = True
    render(cam, pc, Pipe, bg, scaling_modifier=0.7)                   # covariance precomputed in Python
    Pipe.compute_cov3D_python = False
    render(cam, pc, Pipe, bg, override_color=torch.rand(P, 3, generator=torch.Generator().manual_seed(3)))

    flat = {"n_calls": np.array(len(CALLS)), "cam_R": Rc, "cam_T": Tc, "cam_fov": np.array([fovx, fovy]), "WH": np.array([W, H])}
    for i, c in enumerate(CALLS):
        flat[f"c{i}_forward_order"] = np.array(c["forward_order"])
        for grp in ("settings", "forward"):
            for k, d in c[grp].items():
                key = f"c{i}_{grp}_{k}"
                if d is None:
                    flat[key + "_none"] = np.array(1)
                elif isinstance(d, dict):
                    flat[key + "_values"] = d["values"]
                    flat[key + "_meta"] = np.array([str(d["shape"]), str(d["stride"]), d["dtype"], str(d["requires_grad"]),
                                                    str(d["is_parameter"]), str(d["contiguous"])])
                else:
                    flat[key + "_scalar"] = np.array(d)
    np.savez_compressed(OUT, **flat)
    for i, c in enumerate(CALLS):
        print("call", i, {k: (None if v is None else (v["shape"], v["stride"], v["requires_grad"]) if isinstance(v, dict) else v)
                          for k, v in {**c["settings"], **c["forward"]}.items()})
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
