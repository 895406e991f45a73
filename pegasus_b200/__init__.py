"""pegasus_b200 — B200-native (sm_100a) forward renderer for the PEGASUS compose -> rasterize hot path.

Public surface:
    GaussianRasterizationSettings, GaussianRasterizer   drop-in for diff_gaussian_rasterization
    ComposedScene                                        env + posed objects resident in HBM, fused K+3 passes
    Camera                                               view / projection tensors (reference Camera semantics)
    render_rgb_and_depth, render_silhouette_mask, render_visib_mask, render_semanticsegmentation_mask
                                                         mirrors of src/gs/render.py on a ComposedScene
    DatasetGenerator, BOPDatasetWriter, ObjectMeta       the generate_dataset loop (pegasus.py:247-390) with frames
                                                         in flight, GPU-side packing and the BOP writer
    set_numerics("fast" | "exact")                       compositing numerics of calls that do not name one: "fast" (default;
                                                         MUFU exp, inside the reference tolerances) or "exact" (bit-
                                                         reproducible on the CPU oracle); also PG_NUMERICS
Everything computes in libpegasus_b200.so (C ABI, include/pegasus_b200.h); there is no CPU path.
"""
from ._lib import set_numerics  # noqa: F401
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
from .cameras import Camera, focal2fov, fov2focal  # noqa: F401
from .scene import ComposedScene  # noqa: F401
from .render import (render_rgb_and_depth, render_silhouette_mask, render_visib_mask,  # noqa: F401
                     render_semanticsegmentation_mask, render_frame)
from .sh_rotation import generate_pose_packets  # noqa: F401
from .generate import DatasetGenerator  # noqa: F401
from .bop_writer import BOPDatasetWriter, ObjectMeta  # noqa: F401
from . import sweep  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "ComposedScene", "Camera", "focal2fov",
           "fov2focal", "render_rgb_and_depth", "render_silhouette_mask", "render_visib_mask",
           "render_semanticsegmentation_mask", "render_frame", "generate_pose_packets", "DatasetGenerator",
           "BOPDatasetWriter", "ObjectMeta", "set_numerics"]
