"""Fused Adam for the Gaussian map parameters.

Drop-in for the optimizer the reference builds in `Mapping.local_optimize` / `global_optimization`
(SLAM/multiprocess/mapper.py:156,623: `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over the six parameter
groups of `GaussianPointCloud.parametrize`, SLAM/gaussian_pointcloud.py:245-284). Same constructor
arguments, `param_groups`, `state`, `step()` and `zero_grad()`; one kernel launch updates all groups
(torch's default path issues one multi-tensor launch chain per group)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import RtgAdamGroup, check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam: weight_decay / amsgrad are not used by the reference and not implemented")
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        # bucket by (device, betas, eps, step): one launch per bucket (one bucket in the reference's use)
        buckets = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not p.is_cuda or p.dtype != torch.float32:
                    raise TypeError("FusedAdam: parameters must be CUDA float32 tensors")
                if not p.is_contiguous():
                    raise ValueError("FusedAdam: parameters must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                key = (p.device, float(b1), float(b2), float(group["eps"]), int(st["step"]))
                buckets.setdefault(key, []).append((p, p.grad.contiguous(), st, float(group["lr"])))
        for (device, b1, b2, eps, step), items in buckets.items():
            stream = torch.cuda.current_stream(device).cuda_stream
            for i in range(0, len(items), _lib.RTG_ADAM_MAX_GROUPS):
                chunk = items[i:i + _lib.RTG_ADAM_MAX_GROUPS]
                arr = (RtgAdamGroup * len(chunk))()
                for k, (p, g, st, lr) in enumerate(chunk):
                    arr[k].param = p.data_ptr()
                    arr[k].grad = g.data_ptr()
                    arr[k].exp_avg = st["exp_avg"].data_ptr()
                    arr[k].exp_avg_sq = st["exp_avg_sq"].data_ptr()
                    arr[k].numel = p.numel()
                    arr[k].lr = lr
                with torch.cuda.device(device):
                    check(L.rtg_adam_step(arr, len(chunk), b1, b2, eps, step, C.c_void_p(stream)), "rtg_adam_step")
        return loss
