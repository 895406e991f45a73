"""Cross-check against a SECOND, independently written GPU implementation: baseline/upstream_style.cu,
the restatement of the public CUDA forward rasterizer (natural float expressions under nvcc's default
FMA contraction, IEEE expf, CUB scan + 64-bit-key radix sort) that bench.py times as `gpu_baseline`.

What this adds to the oracle tests: the oracle and the product share one explicit operation order;
the baseline lets nvcc contract the upstream-shaped expressions itself and sorts with CUB.  Integer
results (radii, pair count, 64-bit keys, sorted point list, tile ranges) must agree bit for bit —
which checks the "order nvcc contracts the upstream expressions into" assumption of DESIGN.md §2 and
the tile|depth sort order with its index tie-break — and images within north_star's tolerances
(RGB 1e-3 max-abs, depth 1e-4 relative), since expf differs by <= 2 ulp.
"""
import math

import numpy as np
import pytest
import torch

from tests import util
from tests.test_gpu_parity import dev, gpu_forward

pytestmark = pytest.mark.gpu


def baseline_forward(inp, ocam, bg, sh_degree=3):
    import baseline
    d = dev()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(d)
    rast = baseline.UpstreamStyleRasterizer()
    out = rast.forward(t(inp["means3D"]), t(inp["shs"]), t(inp["opacities"]), t(inp["scales"]), t(inp["rotations"]),
                       t(ocam["world_view_transform"]), t(ocam["full_proj_transform"]), t(ocam["camera_center"]),
                       t(np.asarray(bg, np.float32)), ocam["image_width"], ocam["image_height"],
                       math.tan(ocam["FoVx"] * 0.5), math.tan(ocam["FoVy"] * 0.5), sh_degree=sh_degree)
    torch.cuda.synchronize()
    return out, rast


@pytest.mark.parametrize("W,H,bg", [(640, 480, (0, 0, 0)), (333, 217, (1, 1, 1))])
def test_product_matches_upstream_style_baseline(W, H, bg):
    from pegasus_b200 import synth
    from pegasus_b200.scene import export_binning
    env, objs = util.small_scene(n_env=30000, n_obj=(4000,))
    sc = util.merged(env, objs)
    inp = util.activated(sc)
    c = synth.orbit_cameras(1, W, H, seed=3001)[0]
    ocam = util.oracle_cam(c)
    P = inp["means3D"].shape[0]

    base, rast = baseline_forward(inp, ocam, bg)
    color, radii, depth, aux = gpu_forward(inp, ocam, bg, reference_lists=True)
    # integers: bit-exact
    np.testing.assert_array_equal(radii.cpu().numpy(), base["radii"][:P].cpu().numpy())
    assert aux["num_rendered"] == base["num_rendered"] > 0
    keys, plist, ranges = export_binning(dev(), P, W, H, aux["pair_capacity"], aux["num_stored"])
    bkeys, bvals, branges = rast.export(W, H)
    np.testing.assert_array_equal(keys, bkeys)      # 64-bit tile|depth keys in sorted order
    np.testing.assert_array_equal(plist, bvals)     # stable order: equal keys keep index order (CUB LSD sort)
    np.testing.assert_array_equal(ranges, branges)
    # floats: north_star tolerances
    bc, bd = base["color"].cpu().numpy(), base["depth"].cpu().numpy()
    assert np.abs(color.cpu().numpy() - bc).max() <= 1e-3
    rel = np.abs(depth.cpu().numpy() - bd) / np.maximum(np.abs(bd), 1e-3)
    assert rel.max() <= 1e-4


def test_reference_frame_runs_k_plus_3_passes():
    """baseline.reference_frame replays the reference's K+3 rasterizer passes on a ComposedScene; the
    merged-scene pass must agree with the fused product frame, the single-object passes with its
    silhouettes (threshold rule of src/gs/render.py:60-63 applied to the baseline's render)."""
    import baseline
    from pegasus_b200 import Camera, ComposedScene, synth
    d = dev()
    env, objs = util.small_scene(n_env=20000, n_obj=(3000, 2500))
    colors = np.asarray([[0.88, 0.6, 0.32], [0.32, 0.6, 0.88]], dtype=np.float32)
    scene = ComposedScene(env, objs, colors, device=d)
    scene.set_poses(synth.static_poses(2, seed=4000))
    c = synth.orbit_cameras(1, 640, 480, seed=3000)[0]
    cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], c["W"], c["H"], device=d)
    bg = torch.zeros(3, device=d)
    ours = scene.render(cam, bg)
    rast = baseline.UpstreamStyleRasterizer()
    passes = baseline.reference_frame(rast, scene, cam, bg)
    torch.cuda.synchronize()
    assert len(passes) == 2 + 3
    assert passes[0]["num_rendered"] == ours["num_rendered"]
    assert (passes[0]["color"] - ours["color"]).abs().max().item() <= 1e-3
    np.testing.assert_array_equal(passes[0]["radii"].cpu().numpy(), ours["radii"].cpu().numpy())
    # passes 1..K render each object alone with its real SH; sizes follow the object table
    assert passes[1]["radii"].numel() == 3000 and passes[2]["radii"].numel() == 2500
    assert passes[3]["radii"].numel() == passes[4]["radii"].numel() == 5500
