"""CPU test of oracle/mapmerge_oracle.py: Mapping.history_merge restated in numpy against tests/golden/history_merge.npz,
the outputs of the reference's own history_merge method with its unmodified slerp (tests/golden/make_history_merge_golden.py)."""
import os

import numpy as np
import pytest

import helpers
from oracle import mapmerge_oracle as mm

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "history_merge.npz"))


@pytest.mark.parametrize("name", sorted(helpers.HISTORY_MERGE_SIZES))
@pytest.mark.parametrize("max_weight", [0.5, 0.9])
def test_history_merge_matches_reference_golden(name, max_weight):
    hist, cur = helpers.history_merge_inputs(name)
    chk = [float(np.sum(hist[k], dtype=np.float64)) for k in sorted(hist)] + [float(np.sum(cur[k], dtype=np.float64)) for k in sorted(cur)]
    assert np.allclose(GOLD[f"{name}_checksum"], chk, rtol=1e-12, equal_nan=True)   # same inputs as the generator
    out, dot = mm.history_merge(hist, cur, max_weight)
    # lerps: the same fp32 operations in the same order
    for k in ("xyz", "features_dc", "features_rest", "scaling"):
        assert np.array_equal(out[k], GOLD[f"{name}_{max_weight}_{k}"]), k
    # rotation: libm acos / sin and the norm's summation order differ by an ulp or two between numpy and torch; the
    # spherical branch divides by sin(theta_0), so the bound scales with 1 / sin(theta_0) (theta_0 -> pi for dot -> -1)
    g = GOLD[f"{name}_{max_weight}_rotation"]
    assert np.isfinite(g).all()
    assert np.all(np.abs(out["rotation"] - g).max(-1) <= helpers.slerp_tolerance(dot))
    if name == "window":   # the fixture exercises both branches, sign flips and the zero-quaternion rows
        with np.errstate(invalid="ignore"):
            lin = np.isnan(dot) | (np.abs(dot) > 0.9995)
        assert lin.sum() > 100 and (~lin).sum() > 100 and np.isnan(dot).sum() == 3 and (dot < 0).sum() > 20
        # quirk: features and scaling move by the FIRST row's weight, xyz by each row's own
        w = max_weight * hist["confidence"] / (cur["confidence"] + np.float32(1e-6))
        assert w.min() == 0 and 0 < w[0, 0] < 1 and len(np.unique(w)) > 50


def test_history_merge_disabled_by_non_positive_weight():
    hist, cur = helpers.history_merge_inputs("window")
    out, _ = mm.history_merge(hist, cur, 0.0)
    assert out["xyz"] is cur["xyz"] and out["rotation"] is cur["rotation_raw"]   # mapper.py:213-214: returns before touching anything
