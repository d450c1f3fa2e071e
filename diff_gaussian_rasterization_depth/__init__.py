"""Drop-in for the reference's `diff_gaussian_rasterization_depth` package (the module name
`SLAM/render.py:8-13` imports). Same public names; the implementation is rtg_slam_b200 (sm_100a)."""
from rtg_slam_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
