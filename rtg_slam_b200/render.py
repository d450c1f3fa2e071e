"""`Renderer.render` of the reference (SLAM/render.py:21-145) on top of the B200 rasterizer.

Same constructor arguments (an `args` bag with the `renderer_*`, `*_sh_degree`, `color_sigma` keys of
configs/base.yaml:30-31,65-68), same `render(viewpoint_camera, gaussian_data, tile_mask=None)` call and the
same result dictionary. The per-pixel normal map is produced without the boolean-mask indexing of the
reference (render.py:130-133), which forces a host synchronisation (`nonzero`)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


class _NormalMap(torch.autograd.Function):
    """out (3,H,W) = normal[depth_index] where depth_index > -1, zeros elsewhere; backward scatter-adds the pixel
    gradients onto the rows of `normal` (what autograd does for the reference's `normal[index]` expression, so that the
    cosine normal loss of mapper.py:433-446 reaches the rotations)."""

    @staticmethod
    def forward(ctx, normal, depth_index):
        if not normal.is_cuda or normal.dtype != torch.float32:
            raise TypeError("gaussian_data['normal'] must be a CUDA float32 tensor")
        H, W = depth_index.shape[-2:]
        out = torch.empty((3, H, W), dtype=torch.float32, device=normal.device)
        check(_lib.lib().rtg_normal_map(C.c_void_p(normal.detach().contiguous().data_ptr()), C.c_void_p(depth_index.contiguous().data_ptr()),
                                        H, W, C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(normal.device).cuda_stream)),
              "rtg_normal_map")
        ctx.save_for_backward(depth_index)
        ctx.n_rows = normal.shape[0]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (depth_index,) = ctx.saved_tensors
        idx = depth_index.reshape(-1).long()
        valid = idx > -1
        g = torch.zeros((ctx.n_rows, 3), dtype=grad_out.dtype, device=grad_out.device)
        g.index_add_(0, idx[valid], grad_out.reshape(3, -1).t()[valid])
        return g, None


class Renderer:
    def __init__(self, args):
        self.raster_settings = None
        self.rasterizer = None
        self.bg_color = torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda")
        self.renderer_opaque_threshold = args.renderer_opaque_threshold
        self.renderer_normal_threshold = np.cos(np.deg2rad(args.renderer_normal_threshold))
        self.scaling_modifier = 1.0
        self.renderer_depth_threshold = args.renderer_depth_threshold
        self.max_sh_degree = args.max_sh_degree
        self.color_sigma = args.color_sigma
        if args.active_sh_degree < 0:
            self.active_sh_degree = self.max_sh_degree
        else:
            self.active_sh_degree = args.active_sh_degree

    def render(self, viewpoint_camera, gaussian_data, tile_mask=None):
        tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
        tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
        self.raster_settings = GaussianRasterizationSettings(
            image_height=int(viewpoint_camera.image_height),
            image_width=int(viewpoint_camera.image_width),
            tanfovx=tanfovx,
            tanfovy=tanfovy,
            bg=self.bg_color,
            scale_modifier=self.scaling_modifier,
            viewmatrix=viewpoint_camera.world_view_transform,
            projmatrix=viewpoint_camera.full_proj_transform,
            sh_degree=self.active_sh_degree,
            campos=viewpoint_camera.camera_center,
            opaque_threshold=self.renderer_opaque_threshold,
            depth_threshold=self.renderer_depth_threshold,
            normal_threshold=self.renderer_normal_threshold,
            color_sigma=self.color_sigma,
            prefiltered=False,
            debug=False,
            cx=viewpoint_camera.cx,
            cy=viewpoint_camera.cy,
            T_threshold=0.0001,
        )
        self.rasterizer = GaussianRasterizer(raster_settings=self.raster_settings)

        normal = gaussian_data["normal"]
        res = self.rasterizer(
            means3D=gaussian_data["xyz"],
            opacities=gaussian_data["opacity"],
            shs=gaussian_data["shs"],
            colors_precomp=None,
            scales=gaussian_data["scales"],
            rotations=gaussian_data["rotations"],
            cov3D_precomp=None,
            normal_w=normal,
            tile_mask=tile_mask,  # None -> all tiles (render.py:101-108)
        )
        rendered_image, rendered_depth, color_index_map, depth_index_map, color_hit_weight, depth_hit_weight, T_map = res[:7]

        # normal[depth_index_map] where the index is > -1, zeros elsewhere (render.py:130-133): one gather kernel,
        # differentiable w.r.t. the per-Gaussian normals like the reference's indexing expression
        render_normal = _NormalMap.apply(normal, depth_index_map)

        return {
            "render": rendered_image,
            "depth": rendered_depth,
            "normal": render_normal,
            "color_index_map": color_index_map,
            "depth_index_map": depth_index_map,
            "color_hit_weight": color_hit_weight,
            "depth_hit_weight": depth_hit_weight,
            "T_map": T_map,
            "radii": res[7],  # not in the reference's dict: the visibility mapoptim.MapOptimizer.step(radii=...) takes
        }
