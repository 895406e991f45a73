"""Shared builders for the parity tests (CPU oracle side + GPU product side)."""
import math

import numpy as np

import oracle
from pegasus_b200 import synth


def activated(cloud):
    """raw PLY-style cloud -> rasterizer inputs the way render() activates them (numpy, float32)."""
    xyz = cloud["xyz"].astype(np.float32)
    P = xyz.shape[0]
    opacity = (1.0 / (1.0 + np.exp(-cloud["opacity"].astype(np.float32)))).astype(np.float32).reshape(P, 1)
    scales = np.exp(cloud["scaling"].astype(np.float32)).astype(np.float32)
    rot = cloud["rotation"].astype(np.float32)
    rot = (rot / np.sqrt((rot * rot).sum(1, keepdims=True))).astype(np.float32)
    shs = np.concatenate([cloud["features_dc"].reshape(P, 1, 3), cloud["features_rest"].reshape(P, 15, 3)], 1).astype(np.float32)
    return dict(means3D=xyz, opacities=opacity, scales=scales, rotations=rot, shs=shs)


def oracle_cam(c):
    return oracle.camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"])


def oracle_forward(inp, cam, bg, sh_degree=3, **kw):
    W, H = cam["image_width"], cam["image_height"]
    return oracle.rasterize_forward(inp["means3D"], inp["opacities"], cam["world_view_transform"],
                                    cam["full_proj_transform"], cam["camera_center"], bg, W, H,
                                    math.tan(cam["FoVx"] * 0.5), math.tan(cam["FoVy"] * 0.5), sh_degree,
                                    shs=kw.get("shs", inp.get("shs")), colors_precomp=kw.get("colors_precomp"),
                                    scales=kw.get("scales", inp.get("scales")),
                                    rotations=kw.get("rotations", inp.get("rotations")),
                                    cov3D_precomp=kw.get("cov3D_precomp"), scale_modifier=kw.get("scale_modifier", 1.0))


def small_scene(n_env=20000, n_obj=(3000, 2500), seed=0):
    env = synth.make_env(n_env, seed=1000 + seed, extent=0.8, scale_mu=0.012)
    objs = {i + 1: synth.make_object(n, seed=2000 + seed + i) for i, n in enumerate(n_obj)}
    return env, objs


def merged(env, posed_objs):
    sc = {k: env[k] for k in oracle.CLOUD_KEYS}
    for o in posed_objs.values():
        sc = oracle.merge_gaussians(sc, o)
    return sc
