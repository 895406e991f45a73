"""Mirrors of the reference's render helpers (/root/reference/src/gs/render.py:14-129) on a
ComposedScene.  The reference runs K+3 rasterizations per frame and thresholds colour distances on
the CPU; here ONE fused pass (pg_render_composed) produces all of them on the device and the
helpers below only slice / reshape its outputs into the reference's return shapes and dtypes:

    render_rgb_and_depth             -> rgb (H,W,3) cpu f32, depth (H,W,1) cpu f32      (:14-33)
    render_silhouette_mask           -> (H,W,n_colours) float64 0/1                     (:36-65)
    render_visib_mask                -> ((H,W,n_colours) float64 0/1, seg image (H,W,3)) (:68-97)
    render_semanticsegmentation_mask -> (H,W,3) uint8                                   (:100-129)
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .scene import ComposedScene


def render_frame(cam, scene: ComposedScene, bg: torch.Tensor, masks: bool = True, **kw) -> Dict[str, torch.Tensor]:
    """The fused frame (device tensors).  Cached on the scene per (camera, pose version) is left to
    the caller: each helper below accepts a precomputed `frame=`."""
    return scene.render(cam, bg, masks=masks, **kw)


def render_rgb_and_depth(cam, gs_scene: ComposedScene, pipe_settings=None, bg=None, debug=False, frame=None):
    f = frame if frame is not None else render_frame(cam, gs_scene, bg, masks=False)
    rgb_image = f["color"].cpu().permute((1, 2, 0))
    depth_image = f["depth"].cpu().permute((1, 2, 0))
    return rgb_image, depth_image


def render_silhouette_mask(cam, gs_scene: ComposedScene, width=None, height=None, color_set=None,
                           pipe_settings=None, bg=None, frame=None):
    f = frame if frame is not None else render_frame(cam, gs_scene, bg)
    return f["silhouette"].permute((1, 2, 0)).cpu().numpy().astype(np.float64)


def render_visib_mask(cam, gs_scene: ComposedScene, color_set=None, height=None, width=None, pipe_settings=None,
                      bg=None, frame=None):
    f = frame if frame is not None else render_frame(cam, gs_scene, bg)
    masks = f["visible"].permute((1, 2, 0)).cpu().numpy().astype(np.float64)
    seg_image = f["seg_color"].cpu().permute((1, 2, 0))
    return masks, seg_image


def render_semanticsegmentation_mask(cam, gs_scene: ComposedScene, color_set=None, height=None, width=None,
                                     pipe_settings=None, bg=None, debug=False, frame=None):
    f = frame if frame is not None else render_frame(cam, gs_scene, bg)
    return f["sem_seg"].cpu().numpy()
