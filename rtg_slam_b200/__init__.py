"""rtg_slam_b200 -- B200-native (sm_100a) implementation of RTG-SLAM's data-parallel hot path:
the differentiable Gaussian rasterizer with opaque-surfel depth, the optimizer step and the projective
point-to-plane ICP step, behind the reference's operator API. See DESIGN.md / INTEGRATION.md."""

__all__ = ["rasterizer", "render", "icp", "optim", "scene"]
