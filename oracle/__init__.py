"""CPU oracle for the PEGASUS compose -> rasterize hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``pegasus_b200/`` or
``diff_gaussian_rasterization/`` does.

PARITY UNPINNED for the rasterizer proper: the reference's CUDA rasterizer submodule
(meyerls/depth-diff-gaussian-rasterization @ 0062df97) is absent from /root/reference and the
reference has no tests or golden vectors for this path.  The pieces the
reference *does* carry as Python are pinned by tests/golden/*.npz (made by tools/make_golden.py,
which imports the reference's own modules in the build container):
SH basis, build_rotation / covariance, view+projection matrices, semantic colours,
scipy quaternion round trip, pose-schedule algebra.

Layout mirrors the reference:
  rasterizer   -> pegasus_oracle.c           [orc_* via ctypes]
  cameras      -> GSP/utils/graphics_utils.py:38-71, GSP/scene/cameras.py:48-57
  SH           -> GSP/utils/sh_utils.py:57-117
  pose / merge -> src/gs/gaussian_model.py:482-623, src/gs/pegasus_setup.py:160-226
  passes/masks -> src/gs/render.py:14-129, src/utility/graphic_utils.py:40-60
"""
from __future__ import annotations

import colorsys
import ctypes as C
import math
import os

import numpy as np

from . import build as _build

_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or (
            os.path.exists(_build.SRC) and os.path.getmtime(path) < os.path.getmtime(_build.SRC)
        ):
            path = _build.build()
        L = C.CDLL(path)
        L.orc_expf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_cov3d.argtypes = [_f32p, C.c_float, _f32p, _f32p]
        L.orc_preprocess.argtypes = [
            C.c_int, C.c_int, _f32p, C.c_void_p, C.c_float, C.c_void_p, _f32p, C.c_void_p,
            C.c_void_p, C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float,
            _i32p, _f32p, _f32p, _f32p, _f32p, _f32p, _u32p, _i32p]
        L.orc_binning.restype = C.c_uint64
        L.orc_binning.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _f32p, _u32p, _i32p, _u64p,
                                  _u32p, _u32p]
        L.orc_count_pairs.restype = C.c_uint64
        L.orc_count_pairs.argtypes = [C.c_int, _u32p]
        L.orc_composite.argtypes = [C.c_int, C.c_int, _u32p, _u32p, _f32p, _f32p, _f32p, _f32p,
                                    _f32p, _f32p, _f32p, _f32p, _u32p]
        _LIB = L
    return _LIB


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def expf(x: float) -> float:
    return float(lib().orc_expf(C.c_float(x)))


def cov3d(scale, mod, quat):
    out = np.zeros(6, np.float32)
    lib().orc_cov3d(_f32(scale), float(mod), _f32(quat), out)
    return out


# --------------------------------------------------------------------------------------------
# Rasterizer forward (boundary GSP/gaussian_renderer/__init__.py:38-53,87-95)
# --------------------------------------------------------------------------------------------
def preprocess(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
               sh_degree=3, shs=None, colors_precomp=None, scales=None, rotations=None,
               cov3D_precomp=None, scale_modifier=1.0):
    """viewmatrix / projmatrix are given the way the reference passes them: the *transposed*
    4x4 (GSP/scene/cameras.py:54-56); flattened C-order this is the [4*col+row] layout."""
    means3D = _f32(means3D)
    P = means3D.shape[0]
    opac = _f32(opacities).reshape(-1)
    V = _f32(viewmatrix).reshape(16)
    M = _f32(projmatrix).reshape(16)
    cam = _f32(campos).reshape(3)
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    shs_ = None if shs is None else _f32(shs).reshape(P, 48)
    col_ = None if colors_precomp is None else _f32(colors_precomp).reshape(P, 3)
    sc_ = None if scales is None else _f32(scales).reshape(P, 3)
    ro_ = None if rotations is None else _f32(rotations).reshape(P, 4)
    cv_ = None if cov3D_precomp is None else _f32(cov3D_precomp).reshape(P, 6)
    out = dict(
        radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), np.float32), depth=np.zeros(P, np.float32),
        cov3d=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
        rgb=np.zeros((P, 3), np.float32), tiles_touched=np.zeros(P, np.uint32),
        rect=np.zeros((P, 4), np.int32))
    if P:
        lib().orc_preprocess(P, int(sh_degree), means3D, _ptr(sc_), float(scale_modifier), _ptr(ro_),
                             opac, _ptr(shs_), _ptr(cv_), _ptr(col_), V, M, cam, int(W), int(H),
                             float(tanfovx), float(tanfovy), out["radii"], out["xy"], out["depth"],
                             out["cov3d"], out["conic_opacity"], out["rgb"], out["tiles_touched"],
                             out["rect"])
    return out


def binning(pre, W, H):
    P = pre["radii"].shape[0]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = int(pre["tiles_touched"].astype(np.uint64).sum())
    keys = np.zeros(max(R, 1), np.uint64)
    vals = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    if P:
        R2 = lib().orc_binning(P, int(W), int(H), pre["radii"], pre["depth"], pre["tiles_touched"],
                               pre["rect"], keys, vals, ranges)
        assert R2 == R
    return dict(num_rendered=R, keys=keys[:R], point_list=vals[:R], ranges=ranges)


def composite(pre, bins, bg, W, H):
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    pl = bins["point_list"] if bins["num_rendered"] else np.zeros(1, np.uint32)
    lib().orc_composite(int(W), int(H), bins["ranges"].reshape(-1), np.ascontiguousarray(pl),
                        pre["xy"], pre["rgb"], pre["depth"], pre["conic_opacity"], _f32(bg).reshape(3),
                        color, depth, final_T, n_contrib)
    return dict(color=color, depth=depth, final_T=final_T, n_contrib=n_contrib)


def clip_to_tile_rows(pre, row0, row1):
    """Restrict the tile rectangles of a preprocess result to tile rows [row0,row1) (bench sampling)."""
    r = pre["rect"]
    r[:, 1] = np.maximum(r[:, 1], row0)
    r[:, 3] = np.minimum(r[:, 3], row1)
    h = np.maximum(r[:, 3] - r[:, 1], 0)
    tt = (np.maximum(r[:, 2] - r[:, 0], 0) * h).astype(np.uint32)
    tt[pre["radii"] <= 0] = 0
    pre["tiles_touched"][:] = tt
    pre["radii"][tt == 0] = 0
    r[tt == 0] = 0


def rasterize_forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tanfovx,
                      tanfovy, sh_degree=3, shs=None, colors_precomp=None, scales=None,
                      rotations=None, cov3D_precomp=None, scale_modifier=1.0, band=None):
    """Full forward; returns every intermediate the parity tests compare.
    band=(row0,row1): bench sampling only — bin and composite tile rows [row0,row1) (other tiles
    stay background); the per-Gaussian stage still runs on every Gaussian."""
    pre = preprocess(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
                     sh_degree, shs, colors_precomp, scales, rotations, cov3D_precomp, scale_modifier)
    if band is not None:
        clip_to_tile_rows(pre, band[0], band[1])
    bins = binning(pre, W, H)
    img = composite(pre, bins, bg, W, H)
    out = {}
    out.update(pre)
    out.update(bins)
    out.update(img)
    return out


# --------------------------------------------------------------------------------------------
# Cameras (GSP/utils/graphics_utils.py:38-77, GSP/scene/cameras.py:48-57)
# --------------------------------------------------------------------------------------------
def world2view2(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = np.asarray(R).transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    cam_center = C2W[:3, 3]
    cam_center = (cam_center + translate) * scale
    C2W[:3, 3] = cam_center
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def projection_matrix(znear, zfar, fovX, fovY):
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = np.zeros((4, 4), np.float32)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def camera(R, T, FoVx, FoVy, W, H, znear=0.01, zfar=100.0):
    """Returns the tensors a reference ``Camera`` exposes (all float32, transposed storage)."""
    wvt = world2view2(R, T).transpose(1, 0).copy()
    proj = projection_matrix(znear, zfar, FoVx, FoVy).transpose(1, 0).copy()
    full = (wvt.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)
    center = np.linalg.inv(wvt.astype(np.float32))[3, :3].astype(np.float32)
    return dict(world_view_transform=wvt, projection_matrix=proj, full_proj_transform=full,
                camera_center=center, FoVx=FoVx, FoVy=FoVy, image_width=W, image_height=H,
                R=np.asarray(R), T=np.asarray(T))


# --------------------------------------------------------------------------------------------
# SH (GSP/utils/sh_utils.py) and semantic colours (src/utility/graphic_utils.py:40-60)
# --------------------------------------------------------------------------------------------
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
         0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def sh_basis(dirs):
    """(N,3) unit directions -> (N,16) basis values, index order of eval_sh."""
    d = np.asarray(dirs, dtype=np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    B = np.zeros((d.shape[0], 16))
    B[:, 0] = SH_C0
    B[:, 1] = -SH_C1 * y
    B[:, 2] = SH_C1 * z
    B[:, 3] = -SH_C1 * x
    B[:, 4] = SH_C2[0] * xy
    B[:, 5] = SH_C2[1] * yz
    B[:, 6] = SH_C2[2] * (2.0 * zz - xx - yy)
    B[:, 7] = SH_C2[3] * xz
    B[:, 8] = SH_C2[4] * (xx - yy)
    B[:, 9] = SH_C3[0] * y * (3 * xx - yy)
    B[:, 10] = SH_C3[1] * xy * z
    B[:, 11] = SH_C3[2] * y * (4 * zz - xx - yy)
    B[:, 12] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy)
    B[:, 13] = SH_C3[4] * x * (4 * zz - xx - yy)
    B[:, 14] = SH_C3[5] * z * (xx - yy)
    B[:, 15] = SH_C3[6] * x * (xx - 3 * yy)
    return B


def eval_sh(deg, sh, dirs):
    """sh: (N,3,16) like the reference's ``shs_view``; returns (N,3)."""
    n = (deg + 1) ** 2
    B = sh_basis(dirs)[:, :n]
    return np.einsum("nk,nck->nc", B, np.asarray(sh, np.float64)[:, :, :n])


def rgb2sh(rgb):
    return (np.asarray(rgb) - 0.5) / SH_C0


def generate_colors(n, mode="bgr"):
    cols = []
    for i in range(n):
        rgb = colorsys.hls_to_rgb(i / n, 0.6, 0.7)
        if mode == "bgr":
            cols.append((rgb[2], rgb[1], rgb[0]))
        elif mode == "rgb":
            cols.append(tuple(rgb))
        else:
            raise ValueError("Color mode {} is not supported", mode)
    return np.asarray(cols, dtype=np.float32)


# --------------------------------------------------------------------------------------------
# Pose transform (src/gs/gaussian_model.py:482-546) and schedule (src/gs/pegasus_setup.py:160-226)
# --------------------------------------------------------------------------------------------
def build_rotation(q):
    """GSP/utils/general_utils.py:78-99, float32."""
    q = np.asarray(q, np.float32)
    norm = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / norm[:, None]
    R = np.zeros((q.shape[0], 3, 3), np.float32)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r * z)
    R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y)
    R[:, 2, 1] = 2 * (y * z + r * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def apply_transformation_on_xyz(xyz, R, t):
    """x' = R (x - mean) + mean + t   (src/gs/gaussian_model.py:485-497), float32."""
    xyz = np.asarray(xyz, np.float32)
    R = np.asarray(R, np.float32)
    mean = xyz.mean(axis=0, dtype=np.float32)
    new = xyz - mean
    new = (R @ new.T).T
    new = new + mean
    return (new + np.asarray(t, np.float32)).astype(np.float32)


def apply_rotation_on_splats(rot_wxyz, R):
    """src/gs/gaussian_model.py:499-505 — including its scipy round trips."""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_quat(np.asarray(rot_wxyz, np.float32)).as_quat().astype(np.float32)
    splat_R = build_rotation(q)
    rotated = np.asarray(R, np.float32) @ splat_R
    out = np.roll(Rotation.from_matrix(rotated).as_quat(), 1, axis=-1)
    return out.astype(np.float32)


def sh_rotation_matrices(R, n_dirs=256, seed=7):
    """D_1, D_2, D_3 with  Y_l(d) . (D_l c) = Y_l(R^T d) . c  for the 3DGS real-SH basis, i.e. what
    src/gs/gaussian_model.py:507-516 builds with e3nn (e3nn is not installed; SURVEY §8 a-3 derives
    the identity).  Least-squares over random directions, float64."""
    R = np.asarray(R, np.float64)
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_dirs, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    A = sh_basis(d)
    B = sh_basis(d @ R)  # rows: R^T d
    out = []
    for lo, hi in ((1, 4), (4, 9), (9, 16)):
        D = np.linalg.lstsq(A[:, lo:hi], B[:, lo:hi], rcond=None)[0]
        out.append(D)
    return out


def apply_rotation_on_sh(features_rest, R):
    """features_rest: (N,15,3).  c_l' = D_l c_l per channel (src/gs/gaussian_model.py:518-546)."""
    f = np.array(features_rest, np.float32, copy=True)
    D1, D2, D3 = [d.astype(np.float32) for d in sh_rotation_matrices(R)]
    for D, sl in ((D1, slice(0, 3)), (D2, slice(3, 8)), (D3, slice(8, 15))):
        blk = f[:, sl, :]  # n shs rgb
        f[:, sl, :] = np.einsum("ij,njc->nic", D, blk).astype(np.float32)
    return f


def apply_transformation(cloud, R, t, sh_mode="rotate"):
    """GaussianModel.apply_transformation / PegasusSetup.apply_transformation_on_gs
    (src/gs/gaussian_model.py:579-582, src/gs/pegasus_setup.py:195-207) on a dict cloud."""
    out = dict(cloud)
    out["xyz"] = apply_transformation_on_xyz(cloud["xyz"], R, t)
    out["rotation"] = apply_rotation_on_splats(cloud["rotation"], R)
    if sh_mode == "rotate":
        out["features_rest"] = apply_rotation_on_sh(cloud["features_rest"], R)
    return out


def quat_xyzw_to_matrix(q):
    from scipy.spatial.transform import Rotation
    return Rotation.from_quat(np.asarray(q, np.float64)).as_matrix()


def static_pose(trajectory, object_id):
    """src/gs/pegasus_setup.py:209-226: last recorded step of body 1 gives the step index."""
    last = list(trajectory[str(1)].keys())[-1]
    e = trajectory[str(object_id)][str(last)]
    return quat_xyzw_to_matrix(e["q"]).astype(np.float32), np.asarray(e["t"], np.float32)


def dynamic_pose_delta(trajectory, object_id, timestep):
    """src/gs/pegasus_setup.py:178-193."""
    from scipy.spatial.transform import Rotation
    cur = trajectory[str(object_id)][str(timestep)]
    past = trajectory[str(object_id)][str(timestep - 1)]
    t_delta = (np.asarray(cur["t"]) - np.asarray(past["t"])).astype(np.float32)
    q_delta = Rotation.from_quat(np.asarray(cur["q"])) * Rotation.from_quat(np.asarray(past["q"])).inv()
    return q_delta.as_matrix().astype(np.float32), t_delta


# --------------------------------------------------------------------------------------------
# Scene composition and the K+3 render passes (src/gs/render.py, pegasus.py:254-332)
# --------------------------------------------------------------------------------------------
CLOUD_KEYS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")


def merge_gaussians(a, b):
    """src/gs/gaussian_model.py:584-591 (six vstacks; b's Gaussians get the higher indices)."""
    return {k: np.vstack((a[k], b[k])) for k in CLOUD_KEYS}


def mask_points(a, mask):
    return {k: a[k][mask] for k in CLOUD_KEYS}


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def render(cam, cloud, bg, sh_degree=3, scaling_modifier=1.0, activated=False, band=None):
    """``render()`` of GSP/gaussian_renderer/__init__.py:19-103 on a dict cloud holding raw
    (pre-activation) parameters, as a trained PLY does.  activated=True: opacity / scaling / rotation
    already went through sigmoid / exp / normalize (lets a test hand over the exact device values)."""
    xyz = _f32(cloud["xyz"])
    P = xyz.shape[0]
    if activated:
        opacity, scales, rot = _f32(cloud["opacity"]), _f32(cloud["scaling"]), _f32(cloud["rotation"])
    else:
        opacity = _sigmoid(_f32(cloud["opacity"]))
        scales = np.exp(_f32(cloud["scaling"])).astype(np.float32)
        rot = _f32(cloud["rotation"])
        nrm = np.maximum(np.sqrt((rot * rot).sum(axis=1, keepdims=True)), 1e-12).astype(np.float32)
        rot = (rot / nrm).astype(np.float32)
    shs = np.concatenate((_f32(cloud["features_dc"]).reshape(P, 1, 3),
                          _f32(cloud["features_rest"]).reshape(P, 15, 3)), axis=1)
    tanfovx = math.tan(cam["FoVx"] * 0.5)
    tanfovy = math.tan(cam["FoVy"] * 0.5)
    W, H = int(cam["image_width"]), int(cam["image_height"])
    if P == 0:
        # upstream skips the kernels when P == 0: the zero-initialised images come back as they are
        color = np.zeros((3, H, W), np.float32)
        return dict(render=color, depth=np.zeros((1, H, W), np.float32), radii=np.zeros(0, np.int32),
                    final_T=np.ones((H, W), np.float32))
    out = rasterize_forward(xyz, opacity, cam["world_view_transform"], cam["full_proj_transform"],
                            cam["camera_center"], bg, W, H, tanfovx, tanfovy, sh_degree, shs=shs,
                            scales=scales, rotations=rot, scale_modifier=scaling_modifier, band=band)
    return dict(render=out["color"], depth=out["depth"], radii=out["radii"], final_T=out["final_T"],
                raw=out)


def semantic_object(obj, color):
    """pegasus.py:229-231 + src/gs/render.py:50-52: dc = RGB2SH(colour), rest = 0."""
    o = dict(obj)
    n = obj["xyz"].shape[0]
    o["features_dc"] = np.broadcast_to(rgb2sh(np.asarray(color, np.float32)).astype(np.float32),
                                       (n, 1, 3)).reshape(n, 1, 3).copy()
    o["features_rest"] = np.zeros((n, 15, 3), np.float32)
    return o


def empty_like_env(env):
    return {k: env[k][:0] for k in CLOUD_KEYS}


def render_frame_reference(cam, env, objects, color_set, bg, sh_degree=3, activated=False, band=None):
    """One reference frame = K+3 rasterizations (pegasus.py:254-332, src/gs/render.py:14-129).

    objects: dict bullet_id -> posed cloud (insertion order = merge order).
    Returns rgb (H,W,3), depth (H,W,1), silhouette masks (H,W,n_colours), visible masks
    (H,W,n_colours), sem_seg uint8 (H,W,3).
    """
    W, H = int(cam["image_width"]), int(cam["image_height"])
    scene = {k: env[k] for k in CLOUD_KEYS}
    for oid, obj in objects.items():
        scene = merge_gaussians(scene, obj)
    pkg = render(cam, scene, bg, sh_degree, activated=activated, band=band)
    rgb = pkg["render"].transpose(1, 2, 0)
    depth = pkg["depth"].transpose(1, 2, 0)
    n_col = color_set.shape[0]
    # silhouettes: each object alone on the background (src/gs/render.py:36-65)
    sil = np.zeros((H, W, n_col))
    for oid, obj in objects.items():
        c = color_set[oid - 1]
        sc = merge_gaussians(empty_like_env(env), semantic_object(obj, c))
        img = render(cam, sc, bg, sh_degree, activated=activated, band=band)["render"].transpose(1, 2, 0)
        dist = np.linalg.norm(img - c, axis=2)
        sil[dist <= 0.1, oid - 1] = 1
    # visible masks + semantic segmentation: all objects, no environment (src/gs/render.py:68-129)
    sc = empty_like_env(env)
    for oid, obj in objects.items():
        sc = merge_gaussians(sc, semantic_object(obj, color_set[oid - 1]))
    seg = render(cam, sc, bg, sh_degree, activated=activated, band=band)["render"].transpose(1, 2, 0)
    vis = np.zeros((H, W, n_col))
    for ci, c in enumerate(color_set):
        dist = np.linalg.norm(seg - c, axis=2)
        vis[dist <= 0.1, ci] = 1
    sem = (np.ascontiguousarray(seg) * 255).astype("uint8")
    return dict(rgb=rgb, depth=depth, silhouette=sil, visible=vis, sem_seg=sem, seg_float=seg,
                radii=pkg["radii"], raw=pkg.get("raw"))


# --------------------------------------------------------------------------------------------
# Bench support: the K+3 passes split into a per-frame fixed part (compose, activate, per-Gaussian
# stage of every pass) and a per-band part (binning + compositing + mask tests of tile rows).
# --------------------------------------------------------------------------------------------
def _prepare(cam, cloud, sh_degree=3):
    xyz = _f32(cloud["xyz"])
    P = xyz.shape[0]
    opacity = _sigmoid(_f32(cloud["opacity"]))
    scales = np.exp(_f32(cloud["scaling"])).astype(np.float32)
    rot = _f32(cloud["rotation"])
    rot = (rot / np.maximum(np.sqrt((rot * rot).sum(axis=1, keepdims=True)), 1e-12)).astype(np.float32)
    shs = np.concatenate((_f32(cloud["features_dc"]).reshape(P, 1, 3),
                          _f32(cloud["features_rest"]).reshape(P, 15, 3)), axis=1)
    W, H = int(cam["image_width"]), int(cam["image_height"])
    return preprocess(xyz, opacity, cam["world_view_transform"], cam["full_proj_transform"], cam["camera_center"],
                      W, H, math.tan(cam["FoVx"] * 0.5), math.tan(cam["FoVy"] * 0.5), sh_degree, shs=shs,
                      scales=scales, rotations=rot)


def _finish(pre, bg, W, H, band):
    p = dict(pre)
    for k in ("rect", "tiles_touched", "radii"):
        p[k] = pre[k].copy()
    if band is not None:
        clip_to_tile_rows(p, band[0], band[1])
    bins = binning(p, W, H)
    return composite(p, bins, bg, W, H)


def frame_reference_split(cam, env, objects, color_set, bg, bands, sh_degree=3):
    """Same passes as render_frame_reference; returns (seconds of the fixed part, [seconds per band])."""
    import time
    W, H = int(cam["image_width"]), int(cam["image_height"])
    t0 = time.perf_counter()
    scene = {k: env[k] for k in CLOUD_KEYS}
    for oid, obj in objects.items():
        scene = merge_gaussians(scene, obj)
    passes = [("rgb", _prepare(cam, scene, sh_degree), None)]
    for oid, obj in objects.items():
        c = color_set[oid - 1]
        passes.append(("sil", _prepare(cam, merge_gaussians(empty_like_env(env), semantic_object(obj, c)), sh_degree), c))
    for name in ("vis", "sem"):  # the reference renders the objects-only scene twice (src/gs/render.py:86,118)
        sc = empty_like_env(env)
        for oid, obj in objects.items():
            sc = merge_gaussians(sc, semantic_object(obj, color_set[oid - 1]))
        passes.append((name, _prepare(cam, sc, sh_degree), None))
    t_fixed = time.perf_counter() - t0
    t_bands = []
    for band in bands:
        t0 = time.perf_counter()
        y0, y1 = band[0] * 16, min(H, band[1] * 16)
        for name, pre, c in passes:
            img = _finish(pre, bg, W, H, band)["color"][:, y0:y1, :].transpose(1, 2, 0)
            if name == "sil":
                _ = np.linalg.norm(img - c, axis=2) <= 0.1
            elif name == "vis":
                for cc in color_set:
                    _ = np.linalg.norm(img - cc, axis=2) <= 0.1
            elif name == "sem":
                _ = (np.ascontiguousarray(img) * 255).astype("uint8")
        t_bands.append(time.perf_counter() - t0)
    return t_fixed, t_bands
