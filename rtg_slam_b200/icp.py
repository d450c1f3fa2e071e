"""Frame-to-model projective point-to-plane ICP, API of the reference's SLAM/icp.py.

`ICP` and `IcpTracker` keep the constructor arguments, method names, argument order and return values of
the reference (SLAM/icp.py:16-48,357-452); the arithmetic runs in librtg_splat_b200.so: one kernel per
pyramid level build, one kernel per Gauss-Newton iteration (residuals + Jacobians + 27-term reduction +
damped 6x6 solve + exp_se3 + pose update on the device). A whole `predict_pose` issues ~25 launches and
one 72-byte read-back, against ~600 eager ops and >= 45 host synchronisations in the reference."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


_WS = {}


def _workspace(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WS:
        n = _lib.lib().rtg_icp_workspace_bytes(0, 0)
        _WS[idx] = torch.zeros(n, dtype=torch.uint8, device=torch.device("cuda", idx))
    return _WS[idx]


def _map3(t, name):
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 3 or t.shape[-1] != 3:
        raise TypeError(f"{name} must be a CUDA float32 (H, W, 3) tensor")
    return t.contiguous()


def _intrinsics(K):
    """(fx, fy, cx, cy) python floats from a 3x3 tensor / array (one tiny read-back if it lives on the GPU)."""
    if isinstance(K, torch.Tensor):
        K = K.detach().cpu().numpy()
    K = np.asarray(K, dtype=np.float32)
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def point2plane_loss(p_t0, p_t1, n_t0, reduce="mean"):
    loss = ((p_t1 - p_t0) * n_t0).sum(dim=-1)
    return (loss * loss).mean() if reduce == "mean" else (loss * loss).sum()


def build_level(depth, pool, fx, fy, cx, cy):
    """One pyramid level: max-pool by `pool`, back-project with the level intrinsics, Sobel normals.
    Returns (vertex, normal), each (H//pool, W//pool, 3)."""
    L = _lib.lib()
    if not depth.is_cuda or depth.dtype != torch.float32:
        raise TypeError("depth must be a CUDA float32 tensor")
    H, W = depth.shape[:2]
    d = depth.contiguous()
    Hs, Ws = H // pool, W // pool
    vertex = torch.empty((Hs, Ws, 3), dtype=torch.float32, device=d.device)
    normal = torch.empty((Hs, Ws, 3), dtype=torch.float32, device=d.device)
    check(L.rtg_icp_build_level(_p(d), H, W, pool, fx, fy, cx, cy, _p(vertex), _p(normal), _p(_workspace(d.device)), _stream(d.device)),
          "rtg_icp_build_level")
    return vertex, normal


def build_pyramids(depth, K, n_levels):
    """build_vertex_pyramid + build_normal_pyramid (SLAM/utils.py:511-527): index 0 is the coarsest level."""
    fx, fy, cx, cy = K if isinstance(K, tuple) else _intrinsics(K)
    vs, ns = [], []
    for i in range(n_levels):
        pool = 1 << (n_levels - 1 - i)
        s = 1.0 / pool
        v, n = build_level(depth, pool, np.float32(fx) * np.float32(s), np.float32(fy) * np.float32(s),
                           np.float32(cx) * np.float32(s), np.float32(cy) * np.float32(s))
        vs.append(v)
        ns.append(n)
    return vs, ns


class ICP(torch.nn.Module):
    def __init__(self, max_iter=3, damping=1e-6, distance_threshold=0.2, normal_threshold=20, verbose=False):
        super().__init__()
        self.max_iterations = max_iter
        self.distance_threshold = distance_threshold
        self.normal_threshold = np.cos(np.deg2rad(normal_threshold))
        self.damping = damping
        self.verbose = verbose

    def icp(self, pose10, vertex_t0, vertex_t1, normal_t0, normal_t1, K):
        """Same contract as ICP.icp (SLAM/icp.py:33-48): returns (pose10 (4,4) float32 tensor, valid_ratio tensor)."""
        L = _lib.lib()
        v0, v1 = _map3(vertex_t0, "vertex_t0"), _map3(vertex_t1, "vertex_t1")
        n0, n1 = _map3(normal_t0, "normal_t0"), _map3(normal_t1, "normal_t1")
        device = v0.device
        H, W = v0.shape[:2]
        fx, fy, cx, cy = K if isinstance(K, tuple) else _intrinsics(K)
        pose = pose10.detach().to(device=device, dtype=torch.float32).contiguous().clone()
        valid_ratio = torch.zeros((), dtype=torch.float32, device=device)
        check(L.rtg_icp_solve_level(_p(v0), _p(n0), _p(v1), _p(n1), H, W, fx, fy, cx, cy, float(self.distance_threshold),
                                    float(self.normal_threshold), float(self.damping), int(self.max_iterations), _p(pose),
                                    _p(valid_ratio), _p(_workspace(device)), _stream(device)), "rtg_icp_solve_level")
        return pose, valid_ratio


class IcpTracker:
    def __init__(self, args):
        self.icp_trackers = []
        self.icp_downscales = args.icp_downscales
        self.icp_warmup_frames = args.icp_warmup_frames
        self.icp_use_model_depth = args.icp_use_model_depth
        for iters in args.icp_downscale_iters:
            self.icp_trackers.append(ICP(iters, distance_threshold=args.icp_distance_threshold,
                                         normal_threshold=args.icp_normal_threshold, damping=args.icp_damping,
                                         verbose=args.verbose))
        self.icp_sample_distance_threshold = args.icp_sample_distance_threshold
        self.icp_sample_normal_threshold = args.icp_sample_normal_threshold
        self.icp_fail_threshold = args.icp_fail_threshold
        self.normal_pyramid_t0 = None
        self.vertex_pyramid_t0 = None
        self.verbose = args.verbose
        self.K = None
        self._Kf = None
        self.last_p2ploss = None
        self.last_valid_ratio = None

    def _set_K(self, K):
        self.K = K
        self._Kf = _intrinsics(K)

    def update_curr_status(self, depth_t1, K):
        if self.K is None:
            self._set_K(K)
        self.depth_t1 = depth_t1
        self.vertex_pyramid_t1, self.normal_pyramid_t1 = build_pyramids(depth_t1, self._Kf, len(self.icp_downscales))

    def move_last_status(self):
        self.vertex_pyramid_t0 = self.vertex_pyramid_t1
        self.normal_pyramid_t0 = self.normal_pyramid_t1
        self.last_model_depth = self.depth_t1

    def update_last_status(self, frame, render_depth, frame_depth, render_normal, frame_normal):
        """Fills `render_depth` in place from the measured depth where the model disagrees (SLAM/icp.py:397-415)."""
        L = _lib.lib()
        if not (render_depth.is_contiguous() and render_depth.is_cuda and render_depth.dtype == torch.float32):
            raise TypeError("render_depth must be a contiguous CUDA float32 tensor (it is updated in place)")
        H, W = render_depth.shape[:2]
        fd = frame_depth.contiguous()
        rn, fn = _map3(render_normal, "render_normal"), _map3(frame_normal, "frame_normal")
        check(L.rtg_icp_fill_model_depth(_p(render_depth), _p(fd), _p(rn), _p(fn), H, W, float(self.icp_sample_distance_threshold),
                                         float(self.icp_sample_normal_threshold), _stream(render_depth.device)),
              "rtg_icp_fill_model_depth")
        self.last_model_depth = render_depth

    def predict_pose(self, frame):
        K = frame["K"]
        frame_id = frame["frame_id"]
        if self.vertex_pyramid_t0 is None:
            # the reference evaluates point2plane_loss on a None pyramid here and raises; callers only reach
            # this after move_last_status(), so keep the identity answer and report success
            self._set_K(K)
            return np.eye(4), True
        L = _lib.lib()
        levels = len(self.icp_downscales)
        if self.icp_use_model_depth and frame_id >= self.icp_warmup_frames:
            self.vertex_pyramid_t0, self.normal_pyramid_t0 = build_pyramids(self.last_model_depth, self._Kf, levels)
        device = self.vertex_pyramid_t1[0].device
        Kf = _intrinsics(K) if K is not self.K else self._Kf
        pose = torch.eye(4, dtype=torch.float32, device=device)
        valid_ratio = torch.zeros((), dtype=torch.float32, device=device)
        for level in range(levels):
            s = np.float32(self.icp_downscales[level])
            Kl = tuple(float(np.float32(k) * s) for k in Kf)
            # argument swap of the reference: "0" inside icp() is the CURRENT frame (SLAM/icp.py:438-441)
            pose, valid_ratio = self.icp_trackers[level].icp(pose, self.vertex_pyramid_t1[level], self.vertex_pyramid_t0[level],
                                                             self.normal_pyramid_t1[level], self.normal_pyramid_t0[level], Kl)
        out = torch.empty(18, dtype=torch.float32, device=device)
        out[:16] = pose.reshape(-1)
        out[17] = valid_ratio
        v_t0, v_t1, n_t0 = self.vertex_pyramid_t0[-1], self.vertex_pyramid_t1[-1], self.normal_pyramid_t0[-1]
        H, W = v_t0.shape[:2]
        check(L.rtg_icp_point2plane_loss(_p(v_t0.contiguous()), _p(v_t1.contiguous()), _p(n_t0.contiguous()), H, W, _p(pose),
                                         _p(out[16:17]), _p(_workspace(device)), _stream(device)), "rtg_icp_point2plane_loss")
        host = out.cpu().numpy()  # the only read-back of the solve
        pose_t1_t0 = host[:16].reshape(4, 4).astype(np.float32)
        self.last_p2ploss, self.last_valid_ratio = float(host[16]), float(host[17])
        if self.verbose:
            print(self.last_p2ploss, self.last_valid_ratio)
        tracking_success = not (self.last_p2ploss > self.icp_fail_threshold)
        return pose_t1_t0, tracking_success
