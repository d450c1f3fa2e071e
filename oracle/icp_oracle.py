"""CPU restatement of the reference's projective point-to-plane ICP -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.

Restates, in numpy, SLAM/icp.py of the reference (ICP.icp :33-48, compute_residuals_jacobian :52-104,
compute_jtj/jtr :107-119, lev_mar_H :248-257, least_square_solve/invH :313-333, exp_se3 :271-310,
forward_update_pose :259-268, IcpTracker.predict_pose :417-452, update_last_status :397-415) and the pyramid
helpers of SLAM/utils.py (compute_vertex_map :65-75, feature_gradient :77-98, compute_normal_map :100-122,
build_vertex_pyramid / build_normal_pyramid :511-527).

The arithmetic of the reference lives in PyTorch (pinned 1.13.1 by environment.yaml:8-11; not vendored):
F.grid_sample(nearest, border, align_corners=True) = index round-half-even(clip(((u_n+1)/2)*(W-1), 0, W-1));
nn.MaxPool2d(k) = k x k / stride k maximum, floor output size; F.conv2d with the 3x3 Sobel taps after replicate
padding; torch.inverse of the damped 6x6 normal matrix.

Pinning: tests/golden/icp_*.npz hold outputs of the UNMODIFIED reference file (imported from /root/reference
with stub modules for its missing optional dependencies, torch 2.11 CPU) -- see tests/golden/make_icp_golden.py.
"""
from __future__ import annotations

import numpy as np


def maxpool(depth, k):
    H, W = depth.shape
    Hs, Ws = H // k, W // k
    return depth[: Hs * k, : Ws * k].reshape(Hs, k, Ws, k).max(axis=(1, 3))


def vertex_map(depth, fx, fy, cx, cy):
    H, W = depth.shape
    dt = depth.dtype
    i = np.arange(W, dtype=dt)[None, :]
    j = np.arange(H, dtype=dt)[:, None]
    x = ((i - dt.type(cx)) / dt.type(fx)) * depth
    y = ((j - dt.type(cy)) / dt.type(fy)) * depth
    return np.stack([x, y, depth], axis=-1)


def normal_map(vertex):
    H, W, _ = vertex.shape
    dt = vertex.dtype
    p = np.pad(vertex, ((1, 1), (1, 1), (0, 0)), mode="edge")
    a00, a01, a02 = p[:-2, :-2], p[:-2, 1:-1], p[:-2, 2:]
    a10, a12 = p[1:-1, :-2], p[1:-1, 2:]
    a20, a21, a22 = p[2:, :-2], p[2:, 1:-1], p[2:, 2:]
    dx = -a00 + a02 - 2 * a10 + 2 * a12 - a20 + a22
    dy = -a00 - 2 * a01 - a02 + a20 + 2 * a21 + a22
    n = np.cross(dy.reshape(-1, 3), dx.reshape(-1, 3)).reshape(H, W, 3).astype(dt)
    mag = np.sqrt((n * n).sum(-1, keepdims=True))
    n = n / (mag + dt.type(1e-8))
    d = vertex[..., 2]
    invalid = (d <= d.min()) | (d >= d.max())
    n[invalid] = 0
    return n


def build_pyramids(depth, K, n_levels=3):
    """Index 0 = coarsest. K = (fx, fy, cx, cy) of the full resolution."""
    depth = np.asarray(depth)
    dt = depth.dtype
    vs, ns = [], []
    for i in range(n_levels):
        pool = 1 << (n_levels - 1 - i)
        s = dt.type(1.0 / pool)
        d = maxpool(depth, pool)
        v = vertex_map(d, dt.type(K[0]) * s, dt.type(K[1]) * s, dt.type(K[2]) * s, dt.type(K[3]) * s)
        vs.append(v)
        ns.append(normal_map(v))
    return vs, ns


def exp_se3(xi):
    dt = xi.dtype
    w, v = xi[:3], xi[3:]
    Wh = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=dt)
    W2 = Wh @ Wh
    theta = np.sqrt((w * w).sum())
    I = np.eye(3, dtype=dt)
    if theta <= 1e-8:
        E, J = I, I
    else:
        E = I + Wh * np.sin(theta) / theta + W2 * (1 - np.cos(theta)) / theta**2
        J = I + (1 - np.cos(theta)) / theta**2 * Wh + (theta - np.sin(theta)) / theta**3 * W2
    T = np.eye(4, dtype=dt)
    T[:3, :3] = E
    T[:3, 3] = J @ v
    return T


def residuals_jacobian(pose, v0, v1, n0, n1, K, dist_thr, cos_thr):
    H, W, _ = v0.shape
    dt = v0.dtype
    fx, fy, cx, cy = [dt.type(k) for k in K]
    R, t = pose[:3, :3].astype(dt), pose[:3, 3].astype(dt)
    V = v0.reshape(-1, 3) @ R.T + t
    N = n0.reshape(-1, 3) @ R.T
    with np.errstate(divide="ignore", invalid="ignore"):
        u = (V[:, 0] / V[:, 2]) * fx + cx
        v = (V[:, 1] / V[:, 2]) * fy + cy
        inview = (u > 0) & (u < W - 1) & (v > 0) & (v < H - 1)
        un = u / dt.type((W - 1) / 2) - 1
        vn = v / dt.type((H - 1) / 2) - 1
        ix = ((un + 1) / 2) * dt.type(W - 1)
        iy = ((vn + 1) / 2) * dt.type(H - 1)
    ix = np.nan_to_num(np.clip(ix, 0, W - 1), nan=0.0)
    iy = np.nan_to_num(np.clip(iy, 0, H - 1), nan=0.0)
    xi = np.rint(ix).astype(np.int64)
    yi = np.rint(iy).astype(np.int64)
    r_v = v1[yi, xi]
    r_n = n1[yi, xi]
    mask0 = v0.reshape(-1, 3)[:, 2] > 0
    mask1 = r_v[:, 2] > 0
    diff = V - r_v
    normal_ok = (N * r_n).sum(-1) > cos_thr
    res = (r_n * diff).sum(-1)
    J = np.concatenate([-np.cross(r_n, V), r_n], axis=-1)  # [-(n^T [v]x), n^T]
    with np.errstate(invalid="ignore"):
        occ = ~inview | (np.sqrt((diff * diff).sum(-1)) > dist_thr)
    invalid = occ | ~mask0 | ~mask1 | ~normal_ok
    J[invalid] = 0
    res[invalid] = 0
    return res, J, ~invalid


def icp_level(pose, v0, v1, n0, n1, K, iters, dist_thr=0.1, normal_thr_deg=20.0, damping=1e-4, solve_dtype=None):
    """ICP.icp: `iters` damped Gauss-Newton steps. Returns (pose, valid_ratio)."""
    dt = v0.dtype
    cos_thr = np.cos(np.deg2rad(normal_thr_deg))
    pose = np.asarray(pose, dtype=dt)
    valid = None
    for _ in range(iters):
        res, J, valid = residuals_jacobian(pose, v0, v1, n0, n1, K, dist_thr, cos_thr)
        JtJ = (J.T @ J).astype(dt)
        JtR = (J.T @ res).astype(dt)
        Hm = JtJ + np.trace(JtJ) * dt.type(damping) * np.eye(6, dtype=dt)
        sd = solve_dtype or dt
        xi = (-(np.linalg.inv(Hm.astype(sd)) @ JtR.astype(sd))).astype(dt)
        pose = (exp_se3(xi) @ pose).astype(dt)
    H, W = v0.shape[:2]
    return pose, (valid.sum() / H / W if valid is not None else 0.0)


def predict_pose(depth_prev, depth_curr, K, iters=(5, 5, 5), downscales=(0.25, 0.5, 1.0), dist_thr=0.1, normal_thr_deg=20.0,
                 damping=1e-4, fail_threshold=0.02):
    """IcpTracker.predict_pose for two depth frames: returns (pose 4x4, p2ploss, valid_ratio, success)."""
    dt = np.asarray(depth_curr).dtype
    v_t0, n_t0 = build_pyramids(depth_prev, K, len(downscales))
    v_t1, n_t1 = build_pyramids(depth_curr, K, len(downscales))
    pose = np.eye(4, dtype=dt)
    vr = 0.0
    for lvl, s in enumerate(downscales):
        Kl = tuple(dt.type(k) * dt.type(s) for k in K)
        # argument swap of the reference: "0" inside icp() is the current frame (icp.py:438-441)
        pose, vr = icp_level(pose, v_t1[lvl], v_t0[lvl], n_t1[lvl], n_t0[lvl], Kl, iters[lvl], dist_thr, normal_thr_deg, damping)
    p = v_t1[-1].reshape(-1, 3) @ pose[:3, :3].T + pose[:3, 3]
    l = ((p - v_t0[-1].reshape(-1, 3)) * n_t0[-1].reshape(-1, 3)).sum(-1)
    loss = float((l * l).mean())
    return pose, loss, float(vr), not (loss > fail_threshold)


def fill_model_depth(render_depth, frame_depth, render_normal, frame_normal, dist_thr=0.01, normal_thr=0.01):
    """IcpTracker.update_last_status (icp.py:397-415); arrays (H,W) / (H,W,3); returns the filled depth."""
    na = np.maximum(np.sqrt((render_normal**2).sum(-1)), 1e-8)
    nb = np.maximum(np.sqrt((frame_normal**2).sum(-1)), 1e-8)
    cos = (render_normal * frame_normal).sum(-1) / (na * nb)
    mask = ((np.abs(render_depth - frame_depth) > dist_thr) | (render_depth == 0) | ((1 - cos) > normal_thr)) & (frame_depth > 0)
    out = render_depth.copy()
    out[mask] = frame_depth[mask]
    return out
