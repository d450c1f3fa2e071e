"""CPU checks of two algorithms the kernels rely on, restated index-for-index in numpy (the CUDA code itself is tested
on the GPU): the register / shuffle / shared-memory bitonic network of csrc/binning.cu::sort_in_registers, and the exact
rectangle cut-off test of csrc/common.cuh::rect_below_cutoff (two line minimisations) against brute force."""
import numpy as np
import pytest


def sort_network(keys_in, K, NT=256):
    """Thread t holds positions t*K .. t*K+K-1; sub-stages with distance j >= 32K go through 'shared memory' (partner
    thread t ^ j/K), K <= j < 32K through 'shuffles' (partner lane), j < K inside the thread -- same predicates as the kernel."""
    n = len(keys_in)
    INF = np.uint64(2 ** 64 - 1)
    key = np.full((NT, K), INF, dtype=np.uint64)
    for r in range(K):                       # striped (coalesced) load
        i = r * NT + np.arange(NT)
        key[i < n, r] = keys_in[i[i < n]]
    Nn = 1
    while Nn < n:
        Nn <<= 1
    if K > 1:
        Nn = NT * K                          # real keys occupy every position: the full network is needed
    t = np.arange(NT)
    base = t * K
    size = 2
    while size <= Nn:
        up_t = (base & size) == 0
        j = size >> 1
        while j >= K and j > 0:              # cross-thread sub-stages (shared memory or shuffle: same arithmetic)
            tj = j // K
            keep_min = ((t & tj) == 0) == up_t
            o = key[t ^ tj, :]
            key = np.where((o < key) == keep_min[:, None], o, key)
            j >>= 1
        jj = K // 2
        while jj > 0:                        # in-register sub-stages
            if jj < size:
                for r in range(K):
                    if (r & jj) == 0:
                        up = ((base + r) & size) == 0
                        x, y = key[:, r].copy(), key[:, r | jj].copy()
                        sw = (x > y) == up
                        key[:, r] = np.where(sw, y, x)
                        key[:, r | jj] = np.where(sw, x, y)
            jj >>= 1
        size <<= 1
    return key.reshape(-1)[:n]


@pytest.mark.parametrize("K,sizes", [(1, [1, 2, 31, 33, 100, 255, 256]), (2, [257, 400, 512]), (4, [513, 904, 1024]),
                                     (8, [1025, 1500, 2048]), (16, [2049, 4000, 4096])])
def test_register_bitonic_network_sorts(K, sizes):
    rng = np.random.default_rng(K)
    for n in sizes:
        depth = rng.integers(0, 50, size=n).astype(np.uint64)          # many equal depths, as in a planar scene
        ids = rng.permutation(n).astype(np.uint64)
        keys = (depth << np.uint64(32)) | ids                          # unique keys: (depth bits, Gaussian id)
        assert np.array_equal(sort_network(keys, K), np.sort(keys)), (K, n)


def rect_min_two_lines(gx, gy, a, b, c, x0, x1, y0, y1):
    dxl, dxh, dyl, dyh = gx - x1, gx - x0, gy - y1, gy - y0
    q = lambda dx, dy: b * dx * dy + 0.5 * (a * dx * dx + c * dy * dy)
    cl = lambda v, lo, hi: np.minimum(hi, np.maximum(lo, v))
    dxn, dyn = cl(0.0, dxl, dxh), cl(0.0, dyl, dyh)
    return np.minimum(q(dxn, cl(-b / c * dxn, dyl, dyh)), q(cl(-b / a * dyn, dxl, dxh), dyn))


def test_two_line_minimisation_is_the_rectangle_minimum():
    rng = np.random.default_rng(0)
    N = 50_000
    l1, l2, th = 10 ** rng.uniform(-3, 1, N), 10 ** rng.uniform(-3, 1, N), rng.uniform(0, np.pi, N)
    cs, sn = np.cos(th), np.sin(th)
    a, c, b = l1 * cs * cs + l2 * sn * sn, l1 * sn * sn + l2 * cs * cs, (l1 - l2) * cs * sn   # positive definite conics
    gx, gy = rng.uniform(-40, 60, N), rng.uniform(-40, 60, N)
    x0, y0 = rng.integers(0, 3, N) * 8.0, rng.integers(0, 3, N) * 8.0
    w, h = rng.choice([3.0, 7.0, 15.0], N), rng.choice([3.0, 7.0, 15.0], N)
    got = rect_min_two_lines(gx, gy, a, b, c, x0, x0 + w, y0, y0 + h)
    q = lambda dx, dy: b * dx * dy + 0.5 * (a * dx * dx + c * dy * dy)
    brute = np.full(N, np.inf)
    for u in np.linspace(0, 1, 41):
        for v in np.linspace(0, 1, 41):
            brute = np.minimum(brute, q(gx - (x0 + u * w), gy - (y0 + v * h)))
    assert np.all(got <= brute * (1 + 1e-9) + 1e-12)          # never above any point of the rectangle: never culls wrongly
    inside = (gx >= x0) & (gx <= x0 + w) & (gy >= y0) & (gy <= y0 + h)
    assert np.all(got[inside] == 0)
    # and it is attained: compare with the minimum over the four edges (exact for a convex quadratic with the centre outside)
    cl = lambda v_, lo, hi: np.minimum(hi, np.maximum(lo, v_))
    dxl, dxh, dyl, dyh = gx - (x0 + w), gx - x0, gy - (y0 + h), gy - y0
    edges = np.minimum.reduce([q(dxl, cl(-b / c * dxl, dyl, dyh)), q(dxh, cl(-b / c * dxh, dyl, dyh)),
                               q(cl(-b / a * dyl, dxl, dxh), dyl), q(cl(-b / a * dyh, dxl, dxh), dyh)])
    assert np.allclose(got[~inside], edges[~inside], rtol=1e-12, atol=0)


def test_rect_cell_without_integer_division():
    # (local + 0.5) * (1 / rw) truncates to local // rw for every rectangle the tile grid allows (common.cuh::rect_cell)
    for rw in range(1, 130):
        local = np.arange(0, rw * 80, dtype=np.float32)
        inv = np.float32(1.0) / np.float32(rw)
        dy = ((local + np.float32(0.5)) * inv).astype(np.int32)
        assert np.array_equal(dy, (np.arange(0, rw * 80) // rw).astype(np.int32)), rw


# ----------------------------------------------------------------------------- csrc/ssim.cu, tile for tile
from oracle.ssim_oracle import ssim_tiled, ssim_window as _ssim_window  # noqa: E402  (the restatement lives with the oracles)


@pytest.mark.parametrize("shape", [(3, 37, 53), (1, 16, 16), (2, 5, 40), (1, 33, 7)])
def test_ssim_tiles_and_derivative_maps_match_the_reference_formula(shape):
    """The separable, tiled evaluation of csrc/ssim.cu and its closed-form backward against the reference's expression
    (utils/loss_utils.py:40-100: grouped 11x11 conv2d with zero padding) differentiated by autograd, in float64."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(sum(shape))
    a_np, b_np = rng.uniform(size=shape), rng.uniform(size=shape)
    b_np[:, : shape[1] // 2] = a_np[:, : shape[1] // 2] + 0.01 * rng.normal(size=(shape[0], shape[1] // 2, shape[2]))  # similar half
    a = torch.tensor(a_np, requires_grad=True)
    b = torch.tensor(b_np)
    g = torch.tensor(_ssim_window().astype(np.float64))
    w = (g[:, None] @ g[None, :])[None, None].expand(shape[0], 1, 11, 11).contiguous()
    conv = lambda x: F.conv2d(x, w, padding=5, groups=shape[0])
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 ** 2, conv(b * b) - mu2 ** 2, conv(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 ** 2 + mu2 ** 2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))
    ref = 1 - m.mean()
    ref.backward()
    got, grad = ssim_tiled(a_np, b_np)
    assert abs(got - float(ref.detach())) < 1e-12
    assert np.abs(grad - a.grad.numpy()).max() < 1e-12 * max(1.0, np.abs(a.grad.numpy()).max()) + 1e-15
