"""Size-independent properties of the forward rasterizer, checked on the oracle (the GPU path is bit-identical to
it, tests/test_gpu_parity.py): what must hold whatever the scene is."""
import numpy as np
import pytest

from pegasus_b200 import synth
from tests import util


@pytest.fixture(scope="module")
def scene():
    env, objs = util.small_scene(n_env=1500, n_obj=(300,), seed=21)
    inp = util.activated(util.merged(env, objs))
    c = synth.orbit_cameras(1, 160, 112, seed=3021)[0]
    cam = util.oracle_cam(c)
    rng = np.random.default_rng(5)
    cols = rng.uniform(0.0, 1.0, size=(inp["means3D"].shape[0], 3)).astype(np.float32)
    return inp, cam, cols


def fwd(inp, cam, bg, cols):
    return util.oracle_forward(inp, cam, np.asarray(bg, np.float32), colors_precomp=cols, shs=None)


def test_background_enters_as_final_transmittance_times_bg(scene):
    inp, cam, cols = scene
    a = fwd(inp, cam, (0, 0, 0), cols)
    b = fwd(inp, cam, (0.25, 0.5, 1.0), cols)
    np.testing.assert_array_equal(a["final_T"], b["final_T"])
    np.testing.assert_array_equal(a["depth"], b["depth"])
    want = a["color"].astype(np.float64) + a["final_T"].astype(np.float64)[None] * np.array([0.25, 0.5, 1.0])[:, None, None]
    np.testing.assert_allclose(b["color"], want, rtol=0, atol=1.2e-7)   # one fma rounding
    assert a["final_T"].min() < 0.1 < a["final_T"].max()                 # covered and uncovered pixels both exist


def test_colour_is_linear_in_the_gaussian_colours(scene):
    inp, cam, cols = scene
    a = fwd(inp, cam, (0, 0, 0), cols)
    h = fwd(inp, cam, (0, 0, 0), cols * np.float32(0.5))                 # powers of two scale exactly
    np.testing.assert_array_equal(h["color"], a["color"] * np.float32(0.5))
    np.testing.assert_array_equal(h["radii"], a["radii"])
    np.testing.assert_array_equal(h["n_contrib"], a["n_contrib"])


def test_fully_transparent_gaussians_change_nothing(scene):
    inp, cam, cols = scene
    a = fwd(inp, cam, (0, 0, 0), cols)
    P = inp["means3D"].shape[0]
    extra = {k: np.concatenate([v, v[: P // 3]]) for k, v in inp.items()}
    extra["opacities"][P:] = 0.0                                         # duplicates with alpha 0 < 1/255
    b = fwd(extra, cam, (0, 0, 0), np.concatenate([cols, 1.0 - cols[: P // 3]]))
    np.testing.assert_array_equal(b["color"], a["color"])
    np.testing.assert_array_equal(b["depth"], a["depth"])
    np.testing.assert_array_equal(b["radii"][:P], a["radii"])
    assert b["num_rendered"] > a["num_rendered"]                          # they ARE binned, they just never blend


def test_order_of_the_input_only_matters_for_equal_depths(scene):
    """Blending order is (depth, index): shuffling the Gaussians changes nothing unless two visible ones share a
    depth exactly (they do not here)."""
    import oracle
    inp, cam, cols = scene
    a = fwd(inp, cam, (0, 0, 0), cols)
    pre = oracle.preprocess(inp["means3D"], inp["opacities"], cam["world_view_transform"], cam["full_proj_transform"],
                            cam["camera_center"], cam["image_width"], cam["image_height"],
                            np.tan(cam["FoVx"] * 0.5), np.tan(cam["FoVy"] * 0.5), 3, colors_precomp=cols,
                            scales=inp["scales"], rotations=inp["rotations"])
    vis = pre["radii"] > 0
    assert np.unique(pre["depth"][vis]).size == int(vis.sum())
    perm = np.random.default_rng(9).permutation(inp["means3D"].shape[0])
    b = fwd({k: v[perm] for k, v in inp.items()}, cam, (0, 0, 0), cols[perm])
    np.testing.assert_array_equal(b["radii"], a["radii"][perm])
    assert b["num_rendered"] == a["num_rendered"]
    np.testing.assert_array_equal(b["color"], a["color"])
    np.testing.assert_array_equal(b["depth"], a["depth"])
