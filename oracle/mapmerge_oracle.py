"""CPU restatement (numpy, fp32) of Mapping.history_merge -- TEST INFRASTRUCTURE, never imported by the product.

* history_merge: SLAM/multiprocess/mapper.py:212-250 -- after an optimisation call the raw parameters are pulled back
  towards their pre-optimisation values (`history_stat`, mapper.py:146-155) with the weight
  max_weight * confidence_before / (confidence_now + 1e-6). Quirk kept: `_features_dc`, `_features_rest` and `_scaling`
  are merged with `history_weight[0]`, i.e. with the weight of the FIRST Gaussian for every row (mapper.py:229,234,239);
  `_xyz` and the rotation use the per-row weight. `_opacity` is not merged.
* slerp: SLAM/utils.py:593-652 -- linear interpolation where |dot| of the normalised quaternions exceeds 0.9995 or is NaN
  (a zero quaternion), spherical otherwise; the interpolation uses the un-normalised inputs.

Pinned to tests/golden/history_merge.npz, produced by executing the reference's own `history_merge` method (compiled unchanged
from its file) with its unmodified `slerp` (tests/golden/make_history_merge_golden.py)."""
import numpy as np

F32 = np.float32


def lerp(start, end, w):
    """torch.lerp: start + w * (end - start) for |w| < 0.5, end - (end - start) * (1 - w) otherwise (ATen/native/Lerp.h)."""
    diff = end - start
    return np.where(np.abs(w) < F32(0.5), start + w * diff, end - diff * (F32(1) - w)).astype(F32)


def slerp(v0, v1, t, dot_threshold=0.9995):
    """SLAM/utils.py:593-652. v0, v1: (P,4) fp32; t: (P,1) fp32. Returns (out, dot)."""
    v0, v1, t = v0.astype(F32), v1.astype(F32), t.astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        n0 = np.sqrt((v0 * v0).sum(-1, dtype=F32))                    # :608-609
        n1 = np.sqrt((v1 * v1).sum(-1, dtype=F32))
        dot = ((v0 / n0[:, None]) * (v1 / n1[:, None])).sum(-1, dtype=F32)   # :611-615
        gotta_lerp = np.isnan(dot) | (np.abs(dot) > F32(dot_threshold))      # :620
        out = np.zeros_like(v0)
        out[gotta_lerp] = lerp(v0, v1, t)[gotta_lerp]                        # :632-635
        theta0 = np.arccos(dot)[:, None]                                     # :640-648
        sin0 = np.sin(theta0)
        theta_t = theta0 * t
        s0 = np.sin(theta0 - theta_t) / sin0
        s1 = np.sin(theta_t) / sin0
        sl = (s0 * v0 + s1 * v1).astype(F32)
    out[~gotta_lerp] = sl[~gotta_lerp]
    return out, dot


def normalize(q):
    """F.normalize(q, dim=-1): q / max(|q|, 1e-12) (rotation_activation, SLAM/gaussian_pointcloud.py:23)."""
    n = np.maximum(np.sqrt((q * q).sum(-1, keepdims=True, dtype=F32)), F32(1e-12))
    return (q / n).astype(F32)


def history_merge(hist, cur, max_weight=0.5):
    """hist: dict confidence (P,1), xyz, features_dc (P,1,3), features_rest (P,15,3), scaling, rotation (= get_rotation before
    the optimisation, i.e. normalised); cur: dict confidence, xyz, features_dc, features_rest, scaling, rotation_raw. Returns
    the merged raw parameters (dict xyz, features_dc, features_rest, scaling, rotation) and the quaternion dots."""
    if max_weight <= 0:                                                       # :213-214
        return {k: cur[k2] for k, k2 in (("xyz", "xyz"), ("features_dc", "features_dc"), ("features_rest", "features_rest"),
                                         ("scaling", "scaling"), ("rotation", "rotation_raw"))}, None
    w = (F32(max_weight) * hist["confidence"].astype(F32)) / (cur["confidence"].astype(F32) + F32(1e-6))   # :215-219, (P,1)
    one = F32(1)
    out = {"xyz": (hist["xyz"] * w + (one - w) * cur["xyz"]).astype(F32)}    # :223-226
    w0 = w[0]                                                                 # :229 -- shape (1,): the first row's weight
    for k in ("features_dc", "features_rest", "scaling"):                     # :228-241
        out[k] = (hist[k] * w0 + (one - w0) * cur[k]).astype(F32)
    out["rotation"], dot = slerp(hist["rotation"], normalize(cur["rotation_raw"]), one - w)   # :242-244
    return out, dot
