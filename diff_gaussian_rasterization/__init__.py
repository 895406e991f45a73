"""Drop-in for the module PEGASUS imports at
submodules/gaussian-splatting-pegasus/gaussian_renderer/__init__.py:14:

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Put the repo root on sys.path ahead of the original package and `render()`, src/gs/render.py and
pegasus.py run unchanged on the sm_100a kernels.
"""
from pegasus_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
