"""CPU restatement of the fused SSIM term -- TEST INFRASTRUCTURE, never imported by the product.

Follows `ssim` / `_ssim` / `create_window` / `gaussian` of the reference (utils/loss_utils.py:40-100; called as
`1 - ssim(render, gt)` by Mapping.loss_update, SLAM/multiprocess/mapper.py:411-415) in the SEPARABLE, TILED form
rtg_slam_b200/csrc/ssim.cu evaluates it in: one "CTA" per 16x16 tile and channel, a 26x26 staged tile with zero padding, two
11-tap passes with the kernels' index arithmetic, the three derivative maps and the closed-form backward. Arithmetic is
float64: this checks the tiling, the halo handling and the derivative, not the rounding.

Pinned to tests/golden/ssim.npz, produced by the reference's unmodified `ssim` (tests/golden/make_ssim_golden.py)."""
import numpy as np


def ssim_window():
    g = np.array([np.float32(np.exp(-((x - 5) ** 2) / (2.0 * 1.5 * 1.5))) for x in range(11)], np.float32)
    s = np.float32(0)
    for v in g:
        s = np.float32(s + v)
    return (g / s).astype(np.float32)


def _ssim_cta(planes, H, W, bx, by, g):
    """Two 11-tap passes over the 26x26 staged tile of each plane with the kernels' index arithmetic: staging loop
    k -> (r, q) = (k // 26, k % 26), zero outside the image; horizontal pass k -> (r, q) = (k // 16, k % 16) over 26 rows;
    vertical pass thread tid -> (tx, ty) = (tid & 15, tid >> 4)."""
    T, R, E = 16, 5, 26
    x0, y0 = bx * T, by * T
    staged = np.zeros((len(planes), E, E), np.float64)
    for k in range(E * E):
        r, q = k // E, k % E
        gy, gx = y0 + r - R, x0 + q - R
        if 0 <= gy < H and 0 <= gx < W:
            staged[:, r, q] = [p[gy, gx] for p in planes]
    h = np.zeros((len(planes), E, T))
    for k in range(E * T):
        r, q = k // T, k % T
        h[:, r, q] = (staged[:, r, q:q + 11] * g).sum(-1)
    out = np.zeros((len(planes), T, T))
    for tid in range(256):
        tx, ty = tid & 15, tid >> 4
        out[:, ty, tx] = (h[:, ty:ty + 11, tx] * g).sum(-1)
    return out, x0, y0


def ssim_tiled(img1, img2):
    """1 - mean(ssim_map) and its gradient w.r.t. img1 the way ssim_fwd_kernel / ssim_final_kernel / ssim_bwd_kernel compute
    them (float64 arithmetic: this checks the tiling, the halo and the derivative maps, not the rounding)."""
    C, H, W = img1.shape
    g = ssim_window().astype(np.float64)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    d = np.zeros((3, C, H, W))
    gx_t, gy_t = (W + 15) // 16, (H + 15) // 16
    partial = np.zeros(C * gy_t * gx_t)
    for c in range(C):
        a, b = img1[c], img2[c]
        for by in range(gy_t):
            for bx in range(gx_t):
                (m1, m2, e11, e22, e12), x0, y0 = _ssim_cta([a, b, a * a, b * b, a * b], H, W, bx, by, g)
                A1, A2 = 2 * m1 * m2 + C1, 2 * (e12 - m1 * m2) + C2
                B1, B2 = m1 * m1 + m2 * m2 + C1, (e11 - m1 * m1) + (e22 - m2 * m2) + C2
                inv = 1 / (B1 * B2)
                S = A1 * A2 * inv
                hh, ww = min(16, H - y0), min(16, W - x0)     # threads with gx < W && gy < H
                sl = (slice(y0, y0 + hh), slice(x0, x0 + ww))
                d[0, c][sl] = (2 * m2 * (A2 - A1) * inv - 2 * m1 * S * (1 / B1 - 1 / B2))[:hh, :ww]
                d[1, c][sl] = (-S / B2)[:hh, :ww]
                d[2, c][sl] = (2 * A1 * inv)[:hh, :ww]
                partial[(c * gy_t + by) * gx_t + bx] = S[:hh, :ww].sum()
    mean = partial.sum() / (C * H * W)
    grad = np.zeros_like(img1)
    scale = -1.0 / (C * H * W)
    for c in range(C):
        for by in range(gy_t):
            for bx in range(gx_t):
                (ca, cb, cc), x0, y0 = _ssim_cta([d[0, c], d[1, c], d[2, c]], H, W, bx, by, g)
                hh, ww = min(16, H - y0), min(16, W - x0)
                sl = (slice(y0, y0 + hh), slice(x0, x0 + ww))
                grad[c][sl] = scale * (ca[:hh, :ww] + 2 * img1[c][sl] * cb[:hh, :ww] + img2[c][sl] * cc[:hh, :ww])
    return 1 - mean, grad
