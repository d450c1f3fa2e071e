"""CPU: oracle/frameprep_oracle.py against the golden outputs of the unmodified reference functions
(tests/golden/frameprep_*.npz, made by tests/golden/make_frameprep_golden.py)."""
import os

import numpy as np
import pytest

import helpers
from oracle import frameprep_oracle as fo

TOL = {"depth_map": 2e-6, "vertex_map_c": 4e-6, "normal_map_c": 1e-4, "confidence_map": 1e-4}


@pytest.mark.parametrize("name", sorted(helpers.FRAMEPREP_CASES))
def test_map_preprocess_matches_reference_golden(name):
    cfg = helpers.FRAMEPREP_CASES[name]
    depth, K = helpers.frameprep_inputs(name)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", f"frameprep_{name}.npz"))
    assert np.isclose(gold["checksum"][0], depth.sum(dtype=np.float64), rtol=1e-12)   # same seeded input
    out = fo.map_preprocess(depth, K, cfg["depth_filter"], cfg["min_depth"], cfg["max_depth"], cfg["thresh"])
    assert np.array_equal(out["invalid_confidence_mask"], gold["invalid_confidence_mask"])
    assert 0.05 < out["invalid_confidence_mask"].mean() < 0.6 and (out["depth_map"] > 0).mean() > 0.3   # a non-trivial case
    for k, tol in TOL.items():
        assert out[k].shape == gold[k].shape and np.abs(out[k] - gold[k]).max() <= tol, k


def test_bilateral_fills_holes_and_ignores_zero_neighbours():
    d = np.full((9, 9), 2.0, np.float32)
    d[4, 4] = 0.0                       # a hole is filled from its neighbours (utils.py:577-588)
    d[0, :] = 0.0
    out = fo.bilateral_filter(d, 2, 2, 2)
    assert abs(out[4, 4] - 2.0) < 1e-6 and abs(out[1, 4] - 2.0) < 1e-6
    assert np.all(fo.bilateral_filter(np.zeros((5, 5), np.float32), 2, 2, 2) == 0)   # weight_sum == 0 -> 0
