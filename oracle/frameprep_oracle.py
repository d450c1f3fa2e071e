"""CPU restatement (numpy, fp32) of the tracker-side frame preprocessing -- TEST INFRASTRUCTURE, never imported by the
product. Follows Tracker.map_preprocess (SLAM/multiprocess/tracker.py:97-132) and its helpers in SLAM/utils.py:
bilateralFilter_torch (:550-589), compute_vertex_map (:65-75), feature_gradient / compute_normal_map (:77-122),
compute_confidence_map (:125-138). Pinned to tests/golden/frameprep_*.npz, produced by those unmodified functions
(tests/golden/make_frameprep_golden.py).
"""
import numpy as np

F = np.float32


def bilateral_filter(depth, radius, sigma_color, sigma_space):
    """utils.py:550-589, taps in the same order."""
    d = np.asarray(depth, F).reshape(depth.shape[0], depth.shape[1])
    h, w = d.shape
    pad = np.zeros((h + 2 * radius, w + 2 * radius), F)
    pad[radius:radius + h, radius:radius + w] = d
    wsum = np.zeros_like(d)
    psum = np.zeros_like(d)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            if i * i + j * j > radius * radius:
                continue
            nb = pad[radius + i:radius + i + h, radius + j:radius + j + w]
            sw = F(-(i * i + j * j) / (2 * sigma_space ** 2))
            cw = -((d - nb) ** 2) / F(2 * sigma_color ** 2)
            wgt = np.exp(sw + cw, dtype=F) * (nb != 0)
            wsum = wsum + wgt
            psum = psum + wgt * nb
    with np.errstate(invalid="ignore", divide="ignore"):
        out = psum / wsum
    out[wsum == 0] = 0
    return out.astype(F)


def vertex_map(depth, K):
    """utils.py:65-75."""
    h, w = depth.shape
    fx, fy, cx, cy = (F(K[0][0]), F(K[1][1]), F(K[0][2]), F(K[1][2]))
    i, j = np.meshgrid(np.arange(w, dtype=F), np.arange(h, dtype=F))
    return np.stack([(i - cx) / fx, (j - cy) / fy, np.ones_like(i)], -1).astype(F) * depth[..., None]


def normal_map(vtx):
    """utils.py:77-122 (Sobel with replicate padding, cross(dy, dx), zero at the depth extremes)."""
    p = np.pad(vtx, ((1, 1), (1, 1), (0, 0)), mode="edge")
    a = lambda dy, dx: p[1 + dy:p.shape[0] - 1 + dy, 1 + dx:p.shape[1] - 1 + dx]
    gx = -a(-1, -1) + a(-1, 1) - F(2) * a(0, -1) + F(2) * a(0, 1) - a(1, -1) + a(1, 1)
    gy = -a(-1, -1) - F(2) * a(-1, 0) - a(-1, 1) + a(1, -1) + F(2) * a(1, 0) + a(1, 1)
    n = np.cross(gy, gx).astype(F)
    mag = np.sqrt((n * n).sum(-1, keepdims=True), dtype=F)
    n = n / (mag + F(1e-8))
    z = vtx[..., 2]
    n[(z <= z.min()) | (z >= z.max())] = 0
    return n.astype(F)


def confidence_map(nrm, K):
    """utils.py:125-138."""
    h, w = nrm.shape[:2]
    fx, fy, cx, cy = (F(K[0][0]), F(K[1][1]), F(K[0][2]), F(K[1][2]))
    xs, ys = np.meshgrid(np.arange(w, dtype=F), np.arange(h, dtype=F))
    proj = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1).astype(F)
    proj = proj / (np.sqrt((proj * proj).sum(-1, keepdims=True)) + F(1e-8))
    nn = np.maximum(np.sqrt((nrm * nrm).sum(-1, keepdims=True)), F(1e-8))
    pn = np.maximum(np.sqrt((proj * proj).sum(-1, keepdims=True)), F(1e-8))
    return np.abs(((nrm / nn) * (proj / pn)).sum(-1, keepdims=True)).astype(F)


def map_preprocess(depth, K, depth_filter, min_depth, max_depth, thresh, radius=5, sigma_color=2, sigma_space=2):
    """tracker.py:104-132 on an (H,W) depth image; returns the tensor entries of frame_map."""
    d = np.asarray(depth, F).reshape(depth.shape[0], depth.shape[1])
    df = bilateral_filter(d, radius, sigma_color, sigma_space) if depth_filter else d.copy()
    df[~((df > min_depth) & (df < max_depth))] = 0
    v = vertex_map(df, K)
    n = normal_map(v)
    c = confidence_map(n, K)
    bad = (n == 0).all(-1) | (c[..., 0] < thresh)
    df[bad] = 0
    n[bad] = 0
    v[bad] = 0
    c[bad] = 0
    return {"depth_map": df[..., None], "vertex_map_c": v, "normal_map_c": n, "confidence_map": c, "invalid_confidence_mask": bad}
