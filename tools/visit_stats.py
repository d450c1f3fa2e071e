"""Work statistics of the compositing kernels on the benchmark scene, computed with the CPU oracle (TEST
INFRASTRUCTURE; nothing here is product code). Reproduces the numbers quoted in DESIGN.md section 3 / 6:

  * how much of every tile list is walked before all its pixels terminate,
  * blended (pixel, Gaussian) pairs per pixel,
  * the ideal number of (warp, entry) visits -- visits in which at least one live pixel of the warp's patch blends --
    for the forward's 8x4 and the backward's 8x8 patches, and the visits the exact per-patch cut-off masks admit.

    python tools/visit_stats.py [--gaussians 1000000] [--tiles 120]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.splat_oracle import OracleRender  # noqa: E402
from rtg_slam_b200 import scene  # noqa: E402


def rect_min_q(gx, gy, a, b, c, x0, x1, y0, y1):
    """Minimum of q = 0.5 (a dx^2 + c dy^2) + b dx dy over the pixel rectangle (two line minimisations, common.cuh)."""
    dxl, dxh, dyl, dyh = gx - x1, gx - x0, gy - y1, gy - y0
    q = lambda dx, dy: b * dx * dy + 0.5 * (a * dx * dx + c * dy * dy)
    cl = lambda v, lo, hi: np.minimum(hi, np.maximum(lo, v))
    dxn, dyn = cl(0.0, dxl, dxh), cl(0.0, dyl, dyh)
    return np.minimum(q(dxn, cl(-b / c * dxn, dyl, dyh)), q(cl(-b / a * dyn, dxl, dxh), dyn))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--tiles", type=int, default=120, help="random sample of tiles for the per-pair statistics")
    args = ap.parse_args()
    cam = scene.make_camera("replica")
    g = scene.surfel_room(args.gaussians, seed=2024)
    o = OracleRender(cam, g, precision="f32", nthreads=os.cpu_count() or 1)
    _, nc = o.image_state()
    pl, rg = o.binning()
    geo = o.geom()
    xy, co = geo["xy"].astype(np.float64), geo["conic_opacity"].astype(np.float64)
    H, W = nc.shape
    th, tw = cam.tile_grid
    L = (rg[:, 1] - rg[:, 0]).reshape(th, tw)
    ncp = np.zeros((th * 16, tw * 16), np.int64)
    ncp[:H, :W] = nc
    tmax = ncp.reshape(th, 16, tw, 16).max(axis=(1, 3))
    print(f"instances (reference rectangle rule) {L.sum()}, mean tile list {L.mean():.0f}")
    print(f"walked before every pixel of the tile has terminated: {tmax.sum()} = {tmax.sum() / L.sum():.1%} of the lists")
    print(f"mean position of the last colour contribution per pixel: {nc.mean():.0f}")
    rng = np.random.default_rng(0)
    tiles = rng.choice(th * tw, min(args.tiles, th * tw), replace=False)
    blends = ideal84 = ideal88 = mask84 = mask88 = 0
    for t in tiles:
        ty, tx = divmod(int(t), tw)
        ys, xs = np.mgrid[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]
        ncl = np.where((ys < H) & (xs < W), ncp[ys, xs], 0)
        m = int(ncl.max())
        if m == 0:
            continue
        ids = pl[rg[t, 0]:rg[t, 1]][:m]
        gx, gy = xy[ids, 0], xy[ids, 1]
        a, b, c, op = co[ids, 0], co[ids, 1], co[ids, 2], co[ids, 3]
        dx, dy = gx[None, None, :] - xs[:, :, None], gy[None, None, :] - ys[:, :, None]
        power = -0.5 * (a * dx * dx + c * dy * dy) - b * dx * dy
        alpha = np.minimum(0.99, op * np.exp(power))
        pos = np.arange(m)
        ok = (power <= 0) & (alpha >= 1 / 255) & (pos[None, None, :] < ncl[:, :, None])
        blends += ok.sum()
        ideal84 += ok.reshape(4, 4, 2, 8, m).any(axis=(1, 3)).sum()
        ideal88 += ok.reshape(2, 8, 2, 8, m).any(axis=(1, 3)).sum()
        qcut = np.log(np.maximum(255 * op, 0.999)) + 0.01
        for k in range(8):
            x0, y0 = tx * 16 + (k & 1) * 8, ty * 16 + (k >> 1) * 4
            pm = ncl[(k >> 1) * 4:(k >> 1) * 4 + 4, (k & 1) * 8:(k & 1) * 8 + 8].max()
            mask84 += ((rect_min_q(gx, gy, a, b, c, x0, x0 + 7, y0, y0 + 3) * 0.9999 <= qcut) & (pos < pm)).sum()
        for k in range(4):
            x0, y0 = tx * 16 + (k & 1) * 8, ty * 16 + (k >> 1) * 8
            pm = ncl[(k >> 1) * 8:(k >> 1) * 8 + 8, (k & 1) * 8:(k & 1) * 8 + 8].max()
            mask88 += ((rect_min_q(gx, gy, a, b, c, x0, x0 + 7, y0, y0 + 7) * 0.9999 <= qcut) & (pos < pm)).sum()
    sc = th * tw / len(tiles)
    print(f"sample of {len(tiles)} tiles, scaled to the frame:")
    print(f"  blended (pixel, Gaussian) pairs {blends * sc / 1e6:.1f} M = {blends / (len(tiles) * 256):.0f} per pixel")
    print(f"  ideal (warp, entry) visits: forward 8x4 {ideal84 * sc / 1e6:.2f} M (lanes busy {blends / (ideal84 * 32):.0%}), "
          f"backward 8x8 {ideal88 * sc / 1e6:.2f} M (lanes busy {blends / (ideal88 * 64):.0%})")
    print(f"  visits admitted by the exact patch masks: forward {mask84 * sc / 1e6:.2f} M, backward {mask88 * sc / 1e6:.2f} M")
    o.close()


if __name__ == "__main__":
    main()
