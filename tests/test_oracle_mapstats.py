"""CPU tests of oracle/mapstats_oracle.py: the tile-mask builders against the golden outputs of the unmodified
reference functions (tests/golden/mapstats_tilemasks.npz), accumulate_gaussian_error against hand-computed cases."""
import os

import numpy as np
import pytest

import helpers
from oracle import mapstats_oracle as mo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mapstats_tilemasks.npz"))


@pytest.mark.parametrize("name", sorted(helpers.MAPSTATS_SIZES))
def test_tile_mask_builders_match_reference_golden(name):
    T, err = helpers.mapstats_inputs(name)
    assert np.allclose(GOLD[f"{name}_checksum"], [T.sum(dtype=np.float64), err.sum(dtype=np.float64)], rtol=1e-12)  # same inputs
    pm = T != 1
    assert np.array_equal(mo.pixelmask2tilemask(pm, 16), GOLD[f"{name}_pix2tile"])
    for ratio in (0.5, 0.1):
        assert np.array_equal(mo.transmission2tilemask(pm, 16, ratio), GOLD[f"{name}_trans_{ratio}"])
    for ratio in (0.4, 0.05):
        mask, kth = mo.colorerror2tilemask(err, 16, ratio)
        gold = GOLD[f"{name}_cerr_{ratio}"]
        assert mask.sum() == gold.sum() == int(mask.size * ratio)
        # identical up to ties at the k-th value (torch.topk's tie order is unspecified)
        diff = mask != gold
        mean = mo.tile_mean(err)
        assert np.all(np.abs(mean[diff] - kth) <= 1e-6 * max(kth, 1e-12))


def test_accumulate_gaussian_error_hand_cases():
    H, W, P = 2, 3, 4
    ci = np.array([0, 0, 1, -1, 7, 3], np.int32)      # -1 and 7 are skipped (outside [0,P))
    di = np.array([2, 2, -1, 2, 0, 4], np.int32)
    ce = np.array([0.5, 0.2, 0.9, 9.0, 9.0, 0.1], np.float32)
    de = np.array([0.1, 0.4, 9.0, 0.05, 0.3, 9.0], np.float32)
    ne = np.array([0.0, 0.6, 9.0, 0.2, 0.0, 9.0], np.float32)
    gc, gd, gn, rs = mo.accumulate_gaussian_error(H, W, P, ce, de, ne, ci, di, 0.3, 0.25, 0.5, True)
    assert np.allclose(gc[:, 0], [0.5, 0.9, 0, 0.1]) and np.allclose(gd[:, 0], [0.3, 0, 0.4, 0]) and np.allclose(gn[:, 0], [0, 0, 0.6, 0])
    # rescale: colour>0.3 -> G0 (0.5), G1 (0.9); depth>0.25 -> G2 (0.4), G0 (0.3); normal>0.5 -> G2 (0.6)
    assert np.array_equal(rs[:, 0], [2, 1, 2, 0])
    gc, gd, gn, rs2 = mo.accumulate_gaussian_error(H, W, P, ce, de, ne, ci, di, 0.3, 0.25, 0.5, False)
    assert np.allclose(gc[:, 0], [0.35, 0.9, 0, 0.1]) and np.allclose(gd[:, 0], [0.3, 0, (0.1 + 0.4 + 0.05) / 3, 0])
    assert np.allclose(gn[:, 0], [0, 0, (0.0 + 0.6 + 0.2) / 3, 0]) and np.array_equal(rs2, rs)
    # empty image / no Gaussians
    z = mo.accumulate_gaussian_error(0, 0, 3, [], [], [], [], [], 0, 0, 0, True)
    assert all(a.shape == (3, 1) and not a.any() for a in z)


def test_color_error_map_zeroes_black_pixels():
    r = np.zeros((3, 2, 2), np.float32)
    g = np.full((3, 2, 2), 0.25, np.float32)
    r[:, 0, 0] = [0.5, 0.25, 0.0]
    e = mo.color_error_map(r, g)
    assert np.allclose(e, [[0.5, 0.0], [0.0, 0.0]])
