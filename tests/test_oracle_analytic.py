"""The oracle against an INDEPENDENT restatement of the published forward rasterizer.

The reference's rasterizer source is absent, so the C oracle (oracle/pegasus_oracle.c) cannot be
pinned to reference outputs for the rasterizer arithmetic ("parity unpinned").  What can be
done on the CPU is to check it against a second derivation that shares no code with it: the equations of
3D Gaussian Splatting as published (EWA projection cov2D = J W Sigma W^T J^T + 0.3 I, conic = cov2D^-1,
radius = ceil(3 sqrt(lambda_max)), 16x16-tile rectangles, alpha = min(0.99, o exp(-1/2 d^T conic d)), skips at
power > 0 and alpha < 1/255, front-to-back blending in (depth, index) order until T (1 - alpha) < 1e-4, depth
accumulated with the same weights), written here in float64 numpy with matrices in textbook form instead of
the CUDA code's column-major float32 expression order.  Integer results must agree exactly on these seeded
scenes; images within float32 round-off.
"""
import math

import numpy as np
import pytest

import oracle
from pegasus_b200 import synth
from tests import util

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sh_to_rgb(sh, d):
    """sh (16,3), unit direction d (3,): degree-3 real SH in the 3DGS sign convention, + 0.5, clamped at 0."""
    x, y, z = d
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    r = C0 * sh[0]
    r = r - C1 * y * sh[1] + C1 * z * sh[2] - C1 * x * sh[3]
    r = (r + C2[0] * xy * sh[4] + C2[1] * yz * sh[5] + C2[2] * (2 * zz - xx - yy) * sh[6]
         + C2[3] * xz * sh[7] + C2[4] * (xx - yy) * sh[8])
    r = (r + C3[0] * y * (3 * xx - yy) * sh[9] + C3[1] * xy * z * sh[10] + C3[2] * y * (4 * zz - xx - yy) * sh[11]
         + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[12] + C3[4] * x * (4 * zz - xx - yy) * sh[13]
         + C3[5] * z * (xx - yy) * sh[14] + C3[6] * x * (xx - 3 * yy) * sh[15])
    return np.maximum(r + 0.5, 0.0)


def quat_to_matrix(q):
    r, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def textbook_forward(inp, cam, bg):
    """float64, textbook matrix form.  `cam` is an oracle.camera dict (its matrices are the transposed ones the
    reference passes; transposed back here so that p_view = V @ [p, 1])."""
    W, H = cam["image_width"], cam["image_height"]
    V = np.asarray(cam["world_view_transform"], np.float64).T      # world -> view, column-vector convention
    PV = np.asarray(cam["full_proj_transform"], np.float64).T      # world -> clip
    campos = np.asarray(cam["camera_center"], np.float64)
    tanx, tany = math.tan(cam["FoVx"] * 0.5), math.tan(cam["FoVy"] * 0.5)
    fx, fy = W / (2 * tanx), H / (2 * tany)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    means = inp["means3D"].astype(np.float64)
    P = means.shape[0]
    radii = np.zeros(P, np.int64)
    rect = np.zeros((P, 4), np.int64)
    xy = np.zeros((P, 2)); depth = np.zeros(P); conic = np.zeros((P, 3)); rgb = np.zeros((P, 3))
    op = inp["opacities"].astype(np.float64).reshape(-1)
    Wrot = V[:3, :3]
    for i in range(P):
        ph = np.append(means[i], 1.0)
        t = (V @ ph)[:3]
        if t[2] <= 0.2:
            continue
        clip = PV @ ph
        ndc = clip[:3] / (clip[3] + 1e-7)
        R = quat_to_matrix(inp["rotations"][i].astype(np.float64))
        S = np.diag(inp["scales"][i].astype(np.float64))
        Sigma = R @ S @ S @ R.T
        tx = min(1.3 * tanx, max(-1.3 * tanx, t[0] / t[2])) * t[2]
        ty = min(1.3 * tany, max(-1.3 * tany, t[1] / t[2])) * t[2]
        J = np.array([[fx / t[2], 0.0, -fx * tx / t[2] ** 2], [0.0, fy / t[2], -fy * ty / t[2] ** 2]])
        cov = J @ Wrot @ Sigma @ Wrot.T @ J.T
        a, b, c = cov[0, 0] + 0.3, cov[0, 1], cov[1, 1] + 0.3
        det = a * c - b * b
        if det == 0.0:
            continue
        mid = 0.5 * (a + c)
        lam = mid + math.sqrt(max(0.1, mid * mid - det))
        rad = int(math.ceil(3.0 * math.sqrt(lam)))
        px, py = ((ndc[0] + 1.0) * W - 1.0) * 0.5, ((ndc[1] + 1.0) * H - 1.0) * 0.5
        r0x = min(gx, max(0, int((px - rad) / 16.0))); r1x = min(gx, max(0, int((px + rad + 15) / 16.0)))
        r0y = min(gy, max(0, int((py - rad) / 16.0))); r1y = min(gy, max(0, int((py + rad + 15) / 16.0)))
        if (r1x - r0x) * (r1y - r0y) == 0:
            continue
        radii[i] = rad
        rect[i] = (r0x, r0y, r1x, r1y)
        xy[i] = (px, py); depth[i] = t[2]
        conic[i] = (c / det, -b / det, a / det)
        d = means[i] - campos
        rgb[i] = sh_to_rgb(inp["shs"][i].astype(np.float64), d / np.linalg.norm(d))
    # blending in (depth, index) order; the sort key is the float32 depth, as in the 64-bit tile|depth keys
    order = np.argsort(depth.astype(np.float32), kind="stable")
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    T = np.ones((H, W)); done = np.zeros((H, W), bool)
    C = np.zeros((3, H, W)); D = np.zeros((H, W))
    for i in order:
        if radii[i] == 0:
            continue
        x0, y0, x1, y1 = 16 * rect[i, 0], 16 * rect[i, 1], min(W, 16 * rect[i, 2]), min(H, 16 * rect[i, 3])
        sl = (slice(y0, y1), slice(x0, x1))
        dx, dy = xy[i, 0] - xx[sl], xy[i, 1] - yy[sl]
        power = -0.5 * (conic[i, 0] * dx * dx + conic[i, 2] * dy * dy) - conic[i, 1] * dx * dy
        alpha = np.minimum(0.99, op[i] * np.exp(np.minimum(power, 0.0)))
        blend = (power <= 0.0) & (alpha >= 1.0 / 255.0) & ~done[sl]
        test_T = T[sl] * (1.0 - alpha)
        stop = blend & (test_T < 1e-4)
        done[sl] |= stop
        blend &= ~stop
        w = np.where(blend, alpha * T[sl], 0.0)
        C[(slice(None),) + sl] += rgb[i][:, None, None] * w
        D[sl] += depth[i] * w
        T[sl] = np.where(blend, test_T, T[sl])
    color = C + T[None] * np.asarray(bg, np.float64)[:, None, None]
    return dict(radii=radii, rect=rect, color=color, depth=D[None], final_T=T, order=order)


@pytest.mark.parametrize("seed,W,H,bg", [(0, 96, 64, (0, 0, 0)), (3, 80, 48, (1, 1, 1)), (7, 133, 77, (0.2, 0.4, 0.6))])
def test_oracle_matches_textbook_float64_rasterizer(seed, W, H, bg):
    env, objs = util.small_scene(n_env=500, n_obj=(120,), seed=seed)
    inp = util.activated(util.merged(env, objs))
    c = synth.orbit_cameras(2, W, H, seed=3000 + seed)[1]
    cam = util.oracle_cam(c)
    bgf = np.asarray(bg, np.float32)
    got = util.oracle_forward(inp, cam, bgf)
    ref = textbook_forward(inp, cam, bgf)
    assert int((ref["radii"] > 0).sum()) > 100
    np.testing.assert_array_equal(got["radii"], ref["radii"])
    vis = ref["radii"] > 0
    np.testing.assert_array_equal(got["rect"][vis], ref["rect"][vis])
    assert got["num_rendered"] == int(((ref["rect"][:, 2] - ref["rect"][:, 0]) * (ref["rect"][:, 3] - ref["rect"][:, 1]))[vis].sum())
    # a pixel whose alpha / transmittance lands within float32 round-off of a threshold may take the other branch
    # (one such flip moves a colour by at most alpha = 1/255 and the accumulated depth by at most z / 255)
    cerr = np.abs(got["color"] - ref["color"])
    derr = np.abs(got["depth"] - ref["depth"]) / np.maximum(np.abs(ref["depth"]), 1e-3)
    zmax = float(ref["depth"].max())
    assert np.quantile(cerr, 0.999) <= 2e-5 and cerr.max() <= 5e-3, (np.quantile(cerr, 0.999), cerr.max())
    assert np.quantile(derr, 0.999) <= 2e-5, np.quantile(derr, 0.999)
    assert np.abs(got["depth"] - ref["depth"]).max() <= 5e-3 * max(zmax, 1.0)
    assert np.abs(got["final_T"] - ref["final_T"]).max() <= 5e-3
    assert ref["color"].max() > 0.2 and ref["final_T"].min() < 0.5  # the scene is actually in view


def test_textbook_single_gaussian_known_answer():
    """One isotropic Gaussian on the optical axis: everything has a closed form.  sigma^2 in pixels is
    (f s / z)^2 + 0.3, the centre pixel's alpha is o exp(-1/2 d^2 / sigma^2), colour = alpha * c."""
    W = H = 64
    fov = 2 * math.atan(0.5)                       # tan(fov/2) = 0.5 -> focal = 64 px
    R, T = np.eye(3), np.zeros(3)
    cam = oracle.camera(R, T, fov, fov, W, H)
    z, s, o = 4.0, 0.25, 0.8
    shs = np.zeros((1, 16, 3), np.float32)
    shs[0, 0] = (np.array([0.9, 0.5, 0.1]) - 0.5) / C0
    inp = dict(means3D=np.array([[0.0, 0.0, z]], np.float32), opacities=np.array([[o]], np.float32),
               scales=np.full((1, 3), s, np.float32), rotations=np.array([[1, 0, 0, 0]], np.float32), shs=shs)
    got = util.oracle_forward(inp, cam, np.zeros(3, np.float32))
    var = (64.0 * s / z) ** 2 + 0.3
    assert got["radii"][0] == math.ceil(3 * math.sqrt(var))
    # pixel centres sit at integer coordinates; the projected mean is at ((0 + 1) * 64 - 1) / 2 = 31.5
    for (px, py) in [(31, 31), (32, 31), (40, 35), (20, 44)]:
        d2 = (31.5 - px) ** 2 + (31.5 - py) ** 2
        alpha = min(0.99, o * math.exp(-0.5 * d2 / var))
        want = alpha * np.array([0.9, 0.5, 0.1]) if alpha >= 1 / 255 else np.zeros(3)
        np.testing.assert_allclose(got["color"][:, py, px], want, atol=2e-6)
        np.testing.assert_allclose(got["depth"][0, py, px], z * alpha if alpha >= 1 / 255 else 0.0, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(got["final_T"][py, px], 1 - alpha if alpha >= 1 / 255 else 1.0, atol=2e-6)
