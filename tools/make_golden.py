"""Generate tests/golden/reference_python.npz by RUNNING the reference's own Python.

Runs only in the build container (needs /root/reference); the GPU box never executes this.
Nothing is copied from the reference: modules that import cleanly are imported from where they
lie; functions living in modules whose imports are unavailable here (open3d, e3nn, plyfile,
colmap_wrapper ...) are pulled out of the module's AST and exec'd against stubs, so the code that
runs is still the reference's own text.

What gets pinned (file:line in /root/reference):
  sh_eval        GSP/utils/sh_utils.py:57-112           eval_sh, deg 0..3
  rgb2sh         GSP/utils/sh_utils.py:114-115
  build_rotation GSP/utils/general_utils.py:78-99
  covariance     GSP/utils/general_utils.py:101-110 + strip_symmetric :62-76
                 (src/gs/gaussian_model.py:38-42 build_covariance_from_scaling_rotation)
  camera         GSP/utils/graphics_utils.py:38-77, GSP/scene/cameras.py:54-57
  colors         src/utility/graphic_utils.py:40-60      generate_colors
  pose xyz/quat  src/gs/gaussian_model.py:482-505        apply_transformation_on_xyz, apply_rotation_on_splats
  merge / mask   src/gs/gaussian_model.py:584-623
  pose schedule  src/gs/pegasus_setup.py:160-226 on src/engine/simulation_steps.json (excerpt)
  camera path    src/utility/pose_interpolation.py:58-106
  mask passes    src/gs/render.py:36-129 (orchestration only; rasterizer = the oracle)
"""
import ast
import json
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
GSP = os.path.join(REF, "submodules/gaussian-splatting-pegasus")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, GSP)
sys.path.insert(0, REF)

# ---- make the reference's hard-wired device="cuda" run on CPU --------------------------------
_orig_zeros = torch.zeros
_orig_to = torch.Tensor.to
_orig_asarray = torch.asarray
_orig_ones = torch.ones
_orig_eye = torch.eye


def _strip(kw):
    if "device" in kw:
        kw = dict(kw)
        kw.pop("device")
    return kw


torch.zeros = lambda *a, **k: _orig_zeros(*a, **_strip(k))
torch.ones = lambda *a, **k: _orig_ones(*a, **_strip(k))
torch.eye = lambda *a, **k: _orig_eye(*a, **_strip(k))
torch.asarray = lambda *a, **k: _orig_asarray(*a, **_strip(k))


def _to(self, *a, **k):
    a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
    k = {kk: v for kk, v in k.items() if not (kk == "device" and str(v).startswith("cuda"))}
    if not a and not k:
        return self
    return _orig_to(self, *a, **k)


torch.Tensor.to = _to
torch.Tensor.cuda = lambda self, *a, **k: self


def extract(path, names, glb, cls=None):
    """exec selected top-level functions (or methods of class `cls`) of a reference file."""
    tree = ast.parse(open(path).read())
    out = {}
    body = tree.body
    if cls is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            code = compile(mod, path, "exec")
            ns = {}
            exec(code, glb, ns)
            out[node.name] = ns[node.name]
    missing = set(names) - set(out)
    assert not missing, missing
    return out


def main():
    rng = np.random.default_rng(20241017)
    G = {}

    # ---------------- SH ----------------
    from utils import sh_utils
    N = 64
    dirs = rng.normal(size=(N, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    sh = rng.normal(size=(N, 3, 16))
    G["sh_dirs"] = dirs
    G["sh_coeffs"] = sh
    for deg in range(4):
        G[f"sh_eval_deg{deg}"] = sh_utils.eval_sh(deg, torch.from_numpy(sh), torch.from_numpy(dirs)).numpy()
    rgb = rng.uniform(0, 1, size=(8, 3))
    G["rgb"] = rgb
    G["rgb2sh"] = sh_utils.RGB2SH(torch.from_numpy(rgb)).numpy()
    G["sh_consts"] = np.array([sh_utils.C0, sh_utils.C1] + sh_utils.C2 + sh_utils.C3)

    # ---------------- rotations / covariance ----------------
    from utils import general_utils as gu
    q = rng.normal(size=(32, 4)).astype(np.float32)
    s = np.exp(rng.normal(-4, 0.7, size=(32, 3))).astype(np.float32)
    Rm = gu.build_rotation(torch.from_numpy(q))
    G["quat"] = q
    G["scale"] = s
    G["build_rotation"] = Rm.numpy()
    L = gu.build_scaling_rotation(torch.from_numpy(1.0 * s), torch.from_numpy(q))
    cov = gu.strip_symmetric(L @ L.transpose(1, 2))
    G["covariance"] = cov.numpy()

    # ---------------- cameras ----------------
    from utils import graphics_utils as gr
    from scipy.spatial.transform import Rotation
    cams = []
    for i in range(4):
        Rc = Rotation.from_rotvec(rng.normal(size=3)).as_matrix()
        T = rng.normal(size=3)
        fovx = math.radians(60 + 10 * i)
        W, H = 640, 480
        fovy = gr.focal2fov(gr.fov2focal(fovx, W), H)
        wvt = torch.tensor(gr.getWorld2View2(Rc, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
        proj = gr.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
        full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
        center = wvt.inverse()[3, :3]
        cams.append(dict(R=Rc, T=T, fovx=fovx, fovy=fovy, wvt=wvt.contiguous().numpy(), proj=proj.contiguous().numpy(),
                         full=full.numpy(), center=center.numpy()))
    for k in cams[0]:
        G["cam_" + k] = np.stack([np.asarray(c[k]) for c in cams])

    # ---------------- semantic colours ----------------
    gfx = extract(os.path.join(REF, "src/utility/graphic_utils.py"), ["generate_colors"],
                  dict(colorsys=__import__("colorsys"), torch=torch, Literal=__import__("typing").Literal))
    G["colors_bgr_7"] = gfx["generate_colors"](7).numpy()
    G["colors_rgb_5"] = gfx["generate_colors"](5, "rgb").numpy()

    # ---------------- GaussianModel pose / merge methods ----------------
    gm_glb = dict(torch=torch, np=np, Rotation=Rotation, build_rotation=gu.build_rotation)
    meth = extract(os.path.join(REF, "src/gs/gaussian_model.py"),
                   ["apply_translation_on_xyz", "apply_rotation_on_xyz", "apply_transformation_on_xyz",
                    "apply_rotation_on_splats", "merge_gaussians", "mask_points"], gm_glb, cls="GaussianModel")

    class StubGM:
        optimizer = None
        xyz_gradient_accum = []
        denom = []
        max_radii2D = []

    for k, f in meth.items():
        setattr(StubGM, k, f)

    def make_gm(n, seed):
        r = np.random.default_rng(seed)
        g = StubGM()
        g._xyz = torch.from_numpy(r.normal(size=(n, 3)).astype(np.float32) * 0.1 + r.normal(size=3).astype(np.float32))
        g._features_dc = torch.from_numpy(r.normal(size=(n, 1, 3)).astype(np.float32))
        g._features_rest = torch.from_numpy(r.normal(size=(n, 15, 3)).astype(np.float32) * 0.1)
        g._opacity = torch.from_numpy(r.normal(size=(n, 1)).astype(np.float32))
        g._scaling = torch.from_numpy(r.normal(-5, 0.5, size=(n, 3)).astype(np.float32))
        g._rotation = torch.from_numpy(r.normal(size=(n, 4)).astype(np.float32))
        return g

    g = make_gm(50, 1)
    G["pose_xyz_in"] = g._xyz.numpy().copy()
    G["pose_rot_in"] = g._rotation.numpy().copy()
    Rp = Rotation.from_quat([0.0455, 0.9000, 0.3089, 0.3042]).as_matrix()
    tp = np.array([0.0973, -0.0600, 0.5531])
    T4 = torch.eye(4, dtype=torch.float32)
    T4[:3, :3] = torch.from_numpy(Rp).type(torch.float32)
    T4[:3, 3] = torch.from_numpy(tp).type(torch.float32)
    G["pose_R"] = T4[:3, :3].numpy().copy()
    G["pose_t"] = T4[:3, 3].numpy().copy()
    g.apply_transformation_on_xyz(T=T4)
    g.apply_rotation_on_splats(R=T4[:3, :3])
    G["pose_xyz_out"] = g._xyz.numpy().copy()
    G["pose_rot_out"] = g._rotation.numpy().copy()

    a, b = make_gm(5, 2), make_gm(3, 3)
    a.merge_gaussians(gaussian=b)
    G["merge_xyz"] = a._xyz.numpy().copy()
    G["merge_a_xyz"] = make_gm(5, 2)._xyz.numpy()
    G["merge_b_xyz"] = make_gm(3, 3)._xyz.numpy()
    mask = torch.ones(8, dtype=bool)
    mask[:5] = False
    a.mask_points(mask)
    G["masked_xyz"] = a._xyz.numpy().copy()

    # ---------------- pose schedule on the recorded trajectory ----------------
    traj = json.load(open(os.path.join(REF, "src/engine/simulation_steps.json")))
    # the file nests the per-body steps; find the dict keyed by body id with step dicts
    def find_traj(d):
        if isinstance(d, dict):
            if "1" in d and isinstance(d["1"], dict) and "0" in d["1"] and "q" in d["1"]["0"]:
                return d
            for v in d.values():
                r = find_traj(v)
                if r is not None:
                    return r
        return None
    tr = find_traj(traj)
    steps = sorted(int(k) for k in tr["1"].keys())
    keep = steps[:6] + steps[-2:]
    excerpt = {"1": {str(k): tr["1"][str(k)] for k in keep}}
    G["traj_steps"] = np.array(keep)
    G["traj_t"] = np.array([excerpt["1"][str(k)]["t"] for k in keep])
    G["traj_q"] = np.array([excerpt["1"][str(k)]["q"] for k in keep])

    ps = extract(os.path.join(REF, "src/gs/pegasus_setup.py"),
                 ["static_object_pose", "dynamic_object_pose", "update_object_pose", "apply_transformation_on_gs"],
                 dict(torch=torch, np=np, Rotation=Rotation), cls="PegasusSetup")

    class StubSetup:
        pass

    for k, f in ps.items():
        setattr(StubSetup, k, f)

    class RecObj:
        def __init__(self):
            self.calls = []

        def apply_transformation_on_xyz(self, T):
            self.calls.append(T.numpy().copy())

        def apply_rotation_on_splats(self, R):
            pass

        def apply_rotation_on_sh(self, R):
            pass

    # renumber kept steps 0..n-1 so update_object_pose(timestep) indexes consecutive entries
    seq = {"1": {str(i): excerpt["1"][str(k)] for i, k in enumerate(keep)}}
    su = StubSetup()
    su.object_trajectory = seq
    o = RecObj()
    su.static_object_pose({1: o})
    G["sched_static_T"] = o.calls[0]
    o = RecObj()
    su.dynamic_object_pose({1: o})
    for ts in range(1, 6):
        su.update_object_pose({1: o}, ts)
    G["sched_dynamic_T"] = np.stack(o.calls)

    # ---------------- camera path interpolation ----------------
    from src.utility import pose_interpolation as pi
    p1 = np.eye(4); p2 = np.eye(4)
    p1[:3, :3] = Rotation.from_rotvec([0.1, 0.2, 0.3]).as_matrix(); p1[:3, 3] = [0.5, 0.1, 1.0]
    p2[:3, :3] = Rotation.from_rotvec([0.4, -0.2, 0.9]).as_matrix(); p2[:3, 3] = [0.7, -0.3, 1.2]
    G["interp_p1"] = p1
    G["interp_p2"] = p2
    G["interp_out"] = np.stack([pi.interpolate_pose(t=a, t1=0, pose1=p1, t2=1, pose2=p2)
                                for a in np.linspace(0, 1, 5)[:-1]])

    # ---------------- K+3 mask passes: reference orchestration over the oracle rasterizer -------
    import oracle

    class PipeStub:
        debug = False
        convert_SHs_python = False
        compute_cov3D_python = False

    def render_shim(cam, pc, pipe, bg, scaling_modifier=1.0, override_color=None):
        cloud = dict(xyz=pc._xyz.numpy(), features_dc=pc._features_dc.numpy(),
                     features_rest=pc._features_rest.numpy(), opacity=pc._opacity.numpy(),
                     scaling=pc._scaling.numpy(), rotation=pc._rotation.numpy())
        out = oracle.render(cam, cloud, bg.numpy())
        return {"render": torch.from_numpy(out["render"]), "depth": torch.from_numpy(out["depth"])}

    rn = extract(os.path.join(REF, "src/gs/render.py"),
                 ["render_rgb_and_depth", "render_silhouette_mask", "render_visib_mask",
                  "render_semanticsegmentation_mask"],
                 dict(torch=torch, np=np, copy=__import__("copy"), render=render_shim,
                      RGB2SH=sh_utils.RGB2SH, cv2=None))

    W, H = 96, 64
    fovx = math.radians(70)
    fovy = oracle.focal2fov(W / (2 * math.tan(fovx / 2)), H)
    eye = np.array([0.0, -1.2, 0.8])
    f = -eye / np.linalg.norm(eye)
    r_ = np.cross(f, [0, 0, 1.0]); r_ /= np.linalg.norm(r_)
    d_ = np.cross(f, r_)
    Rc = np.stack([r_, d_, f], axis=1)
    Tc = -Rc.T @ eye
    cam = oracle.camera(Rc, Tc, fovx, fovy, W, H)

    def blob(n, center, spread, seed, scale=-3.6):
        r = np.random.default_rng(seed)
        g = StubGM()
        g._xyz = torch.from_numpy((r.normal(size=(n, 3)) * spread + center).astype(np.float32))
        g._features_dc = torch.from_numpy(r.uniform(-1, 1, size=(n, 1, 3)).astype(np.float32))
        g._features_rest = torch.from_numpy((r.normal(size=(n, 15, 3)) * 0.05).astype(np.float32))
        g._opacity = torch.from_numpy(r.normal(2, 1, size=(n, 1)).astype(np.float32))
        g._scaling = torch.from_numpy(r.normal(scale, 0.3, size=(n, 3)).astype(np.float32))
        g._rotation = torch.from_numpy(r.normal(size=(n, 4)).astype(np.float32))
        return g

    env = blob(400, np.array([0, 0, 0.0]), np.array([0.6, 0.6, 0.02]), 10, scale=-3.0)
    objs = {2: blob(120, np.array([-0.15, 0.0, 0.12]), 0.05, 11), 1: blob(150, np.array([0.1, -0.1, 0.1]), 0.06, 12),
            3: blob(100, np.array([0.05, 0.2, 0.1]), 0.05, 13)}
    colors = gfx["generate_colors"](3)
    for oid, ob in objs.items():
        ob._features_dc_semantics = sh_utils.RGB2SH(colors[oid - 1])
        ob._features_rest_semantics = torch.asarray([0, 0, 0])
    for name, cl in [("env", env)] + [(f"obj{k}", v) for k, v in objs.items()]:
        for attr in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"):
            G[f"mask_{name}{attr}"] = getattr(cl, attr).numpy().copy()
    G["mask_obj_order"] = np.array(list(objs.keys()))
    G["mask_cam_R"] = Rc
    G["mask_cam_T"] = Tc
    G["mask_cam_fov"] = np.array([fovx, fovy])
    G["mask_WH"] = np.array([W, H])
    bg = torch.zeros(3)
    import copy as _copy
    scene = _copy.deepcopy(env)
    for oid in objs:
        scene.merge_gaussians(gaussian=objs[oid])
    rgb_img, depth_img = rn["render_rgb_and_depth"](cam, scene, PipeStub(), bg)
    G["mask_rgb"] = rgb_img.numpy()
    G["mask_depth"] = depth_img.numpy()
    sil = rn["render_silhouette_mask"](cam, objs, env, W, H, colors, PipeStub(), bg)
    vis, seg_image = rn["render_visib_mask"](cam, env, objs, colors, H, W, PipeStub(), bg)
    sem = rn["render_semanticsegmentation_mask"](cam, env, objs, colors, H, W, PipeStub(), bg, False)
    G["mask_silhouette"] = sil.astype(np.uint8)
    G["mask_visible"] = vis.astype(np.uint8)
    G["mask_sem_seg"] = sem
    G["mask_colors"] = colors.numpy()

    out = os.path.join(REPO, "tests/golden/reference_python.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "arrays")


if __name__ == "__main__":
    main()
