"""CPU restatement (numpy) of the map-statistics functions -- TEST INFRASTRUCTURE, never imported by the product.

* accumulate_gaussian_error: submodules/cuda_utils/map_process.cu:33-245 (+ wrapper cuda_utils.cu:17-60). Pinned on the
  GPU against the reference's own extension built unmodified into oracle/_ref/cuda_utils (tests/test_mapstats_gpu.py);
  there is no CPU path of it in the reference, so here the restatement is only checked against hand-computed cases.
* pixelmask2tilemask / transmission2tilemask / colorerror2tilemask: SLAM/utils.py:681-734. Pinned to
  tests/golden/mapstats_tilemasks.npz, produced by the unmodified functions (tests/golden/make_mapstats_golden.py).
"""
import numpy as np


def accumulate_gaussian_error(H, W, P, color_err, depth_err, normal_err, color_index, depth_index, thr_c, thr_d, thr_n, check_max):
    """map_process.cu:33-112 (per-pixel atomics) and :128-151 (mean pass)."""
    ce, de, ne = (np.asarray(a, np.float32).reshape(-1) for a in (color_err, depth_err, normal_err))
    ci, di = (np.asarray(a, np.int64).reshape(-1) for a in (color_index, depth_index))
    gc, gd, gn, rs = (np.zeros(P, np.float32) for _ in range(4))
    cc, cd = np.zeros(P, np.int64), np.zeros(P, np.int64)
    vc = (ci >= 0) & (ci < P)          # :67
    vd = (di >= 0) & (di < P)          # :84
    if check_max:                      # :69-72, :86-90 (initial value 0)
        np.maximum.at(gc, ci[vc], ce[vc])
        np.maximum.at(gd, di[vd], de[vd])
        np.maximum.at(gn, di[vd], ne[vd])
    else:                              # sums in float64, rounded once: the fp32 atomics' order is not defined
        gc = np.bincount(ci[vc], weights=ce[vc].astype(np.float64), minlength=P)
        gd = np.bincount(di[vd], weights=de[vd].astype(np.float64), minlength=P)
        gn = np.bincount(di[vd], weights=ne[vd].astype(np.float64), minlength=P)
    np.add.at(cc, ci[vc], 1)
    np.add.at(cd, di[vd], 1)
    np.add.at(rs, ci[vc & (ce > thr_c)], 1.0)   # :78-81
    np.add.at(rs, di[vd & (de > thr_d)], 1.0)   # :100-103
    np.add.at(rs, di[vd & (ne > thr_n)], 1.0)   # :104-107
    if not check_max:                  # :128-151
        gc = np.where(cc > 0, gc / np.maximum(cc, 1), gc)
        gd = np.where(cd > 0, gd / np.maximum(cd, 1), gd)
        gn = np.where(cd > 0, gn / np.maximum(cd, 1), gn)
    return tuple(np.asarray(a, np.float32).reshape(P, 1) for a in (gc, gd, gn, rs))


def _pool(img, stride, op):
    H, W = img.shape
    th, tw = (H + stride - 1) // stride, (W + stride - 1) // stride
    pad = np.zeros((th * stride, tw * stride), np.float32)      # F.pad(..., value 0), utils.py:697-699
    pad[:H, :W] = img
    blocks = pad.reshape(th, stride, tw, stride)
    if op == "max":
        return blocks.max(axis=(1, 3))
    return (blocks.sum(axis=(1, 3), dtype=np.float64) / (stride * stride)).astype(np.float32)


def pixelmask2tilemask(pixelmask, stride):
    """utils.py:681-692."""
    return _pool(np.asarray(pixelmask, np.float32), stride, "max").astype(np.int32)


def transmission2tilemask(pixelmask, stride, tile_mask_ratio=0.5):
    """utils.py:695-705."""
    return (_pool(np.asarray(pixelmask, np.float32), stride, "avg") > tile_mask_ratio).astype(np.int32)


def tile_mean(img, stride=16):
    return _pool(np.asarray(img, np.float32), stride, "avg")


def colorerror2tilemask(color_error, stride, top_ratio=0.4):
    """utils.py:708-734. Returns (mask float32, kth_value): tiles whose mean equals kth_value are ties of the top-k."""
    mean = tile_mean(color_error, stride)
    k = int(mean.size * top_ratio)
    order = np.argsort(-mean.reshape(-1), kind="stable")
    mask = np.zeros(mean.size, np.float32)
    mask[order[:k]] = 1
    kth = float(mean.reshape(-1)[order[k - 1]]) if k > 0 else np.inf
    return mask.reshape(mean.shape), kth


def color_error_map(render, gt):
    """mapper.py:481-487 on (3,H,W) arrays."""
    r, g = np.asarray(render, np.float32), np.asarray(gt, np.float32)
    err = np.abs(r - g).sum(axis=0, dtype=np.float32)
    err[r.sum(axis=0, dtype=np.float32) == 0] = 0
    return err
