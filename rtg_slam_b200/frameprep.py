"""Tracker-side preprocessing of an incoming RGB-D frame.

`map_preprocess_maps` is the tensor part of `Tracker.map_preprocess` (SLAM/multiprocess/tracker.py:97-132): bilateral
depth filter, valid-range mask, vertex / normal / confidence maps and the invalid-confidence masking -- three kernel
launches instead of the reference's ~250 eager ones when `depth_filter` is on. `bilateralFilter_torch`,
`compute_vertex_map`, `compute_normal_map` and `compute_confidence_map` keep the names and argument meaning of
SLAM/utils.py:65-138,550-589 for callers that use them one by one.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check

_WS = {}


def _ws(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WS:
        _WS[idx] = torch.zeros(_lib.lib().rtg_icp_workspace_bytes(1, 1), dtype=torch.uint8, device=torch.device("cuda", idx))
    return _WS[idx]


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _run(depth, K, depth_filter, radius, sigma_color, sigma_space, min_depth, max_depth, thresh):
    if not depth.is_cuda or depth.dtype != torch.float32:
        raise TypeError("depth must be a CUDA float32 tensor")
    H, W = depth.shape[:2]
    d = depth.reshape(H, W).contiguous()
    dev = d.device
    fx, fy, cx, cy = (float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
    out_d = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    vtx = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    nrm = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    conf = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    bad = torch.empty((H, W), dtype=torch.uint8, device=dev)
    check(_lib.lib().rtg_frame_preprocess(_p(d), H, W, 1 if depth_filter else 0, int(radius), float(sigma_color), float(sigma_space),
                                          float(min_depth), float(max_depth), fx, fy, cx, cy, float(thresh), _p(out_d), _p(vtx),
                                          _p(nrm), _p(conf), _p(bad), _p(_ws(dev)),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rtg_frame_preprocess")
    return out_d, vtx, nrm, conf, bad.view(torch.bool)


def map_preprocess_maps(depth_map, intrinsic, depth_filter=True, min_depth=0.3, max_depth=5.0, invalid_confidence_thresh=0.2,
                        radius=5, sigma_color=2, sigma_space=2):
    """depth_map: (H,W,1) or (H,W) metric depth; intrinsic: 3x3 (tensor, array or nested list). Returns the tensor entries
    of the reference's `frame_map`: depth_map (filtered, range- and confidence-masked), vertex_map_c, normal_map_c,
    confidence_map, invalid_confidence_mask (tracker.py:104-132,152-158; the bilateral parameters are the reference's
    hard-wired 5, 2, 2 of tracker.py:108)."""
    d, v, n, c, bad = _run(depth_map, intrinsic, depth_filter, radius, sigma_color, sigma_space, min_depth, max_depth,
                           invalid_confidence_thresh)
    return {"depth_map": d, "vertex_map_c": v, "normal_map_c": n, "confidence_map": c, "invalid_confidence_mask": bad}


def bilateralFilter_torch(depth, radius, sigma_color, sigma_space):
    """SLAM/utils.py:550-589. Returns (H,W,1)."""
    if not depth.is_cuda or depth.dtype != torch.float32:
        raise TypeError("depth must be a CUDA float32 tensor")
    H, W = depth.shape[:2]
    d = depth.reshape(H, W).contiguous()
    out = torch.empty((H, W, 1), dtype=torch.float32, device=d.device)
    inf = float("inf")
    check(_lib.lib().rtg_frame_preprocess(_p(d), H, W, 1, int(radius), float(sigma_color), float(sigma_space), -inf, inf, 1.0, 1.0,
                                          0.0, 0.0, 0.0, _p(out), None, None, None, None, None,
                                          C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)), "rtg_frame_preprocess")
    return out
