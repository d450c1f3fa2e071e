"""Fused mapping loss: masked L1 on the rendered colour + L1 on the rendered depth, with gradients.

Drop-in for the colour / depth terms of `Mapping.loss_update` (SLAM/multiprocess/mapper.py:402-431,444-451; `l1_loss` of
utils/loss_utils.py:27-31). The reference evaluates them with ~25 eager kernels (boolean-mask compactions included) and
autograd builds the same number again in the backward; here two kernels produce the loss *and* dL/d(colour), dL/d(depth),
which autograd hands straight to the rasterizer backward."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check

_WS = {}


def _ws(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WS:
        _WS[idx] = torch.zeros(_lib.lib().rtg_loss_workspace_bytes(), dtype=torch.uint8, device=torch.device("cuda", idx))
    return _WS[idx]


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _FusedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, depth, depth_index, gt_color, gt_depth, render_mask, color_weight, depth_weight, depth_error_max):
        L = _lib.lib()
        dev = render.device
        for name, t in (("render", render), ("depth", depth), ("gt_color", gt_color), ("gt_depth", gt_depth)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise TypeError(f"{name} must be a CUDA float32 tensor")
        if depth_index.dtype != torch.int32:
            raise TypeError("depth_index must be int32 (the rasterizer's depth_index_map)")
        _, H, W = render.shape
        render, depth, depth_index = render.contiguous(), depth.contiguous(), depth_index.contiguous()
        gt_color, gt_depth = gt_color.contiguous(), gt_depth.contiguous()
        if gt_color.shape == (H, W, 3):
            channels_last = 1
        elif gt_color.shape == (3, H, W):
            channels_last = 0
        else:
            raise ValueError("gt_color must be (H,W,3) or (3,H,W)")
        if gt_depth.numel() != H * W:
            raise ValueError("gt_depth must have H*W elements")
        mask = None
        if render_mask is not None:
            mask = render_mask.to(torch.uint8).contiguous() if render_mask.dtype != torch.uint8 else render_mask.contiguous()
            if mask.numel() != H * W:
                raise ValueError("render_mask must have H*W elements")
        g_color = torch.empty_like(render)
        g_depth = torch.empty_like(depth)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        check(L.rtg_loss_l1(_p(render), _p(depth), _p(depth_index), _p(gt_color), _p(gt_depth), _p(mask), H, W, channels_last,
                            float(color_weight), float(depth_weight), float(depth_error_max), _p(g_color), _p(g_depth), _p(out),
                            _p(_ws(dev)), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rtg_loss_l1")
        ctx.save_for_backward(g_color, g_depth)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, grad_loss, _grad_parts):
        g_color, g_depth = ctx.saved_tensors
        return g_color * grad_loss, g_depth * grad_loss, None, None, None, None, None, None, None


def l1_color_depth_loss(render_output, gt_color, gt_depth, render_mask=None, color_weight=0.8, depth_weight=1.0,
                        depth_error_max=0.1):
    """`render_output` is the dict of Renderer.render. Defaults: color_weight / depth_weight / add_depth_thres of
    configs/base.yaml:76-77,51. Returns (loss, parts) with parts = tensor [loss, colour_l1, depth_l1, n_depth]."""
    return _FusedL1.apply(render_output["render"], render_output["depth"], render_output["depth_index_map"], gt_color, gt_depth,
                          render_mask, color_weight, depth_weight, depth_error_max)
