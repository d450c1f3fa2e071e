"""Generates tests/golden/history_merge.npz: the arithmetic of Mapping.history_merge (SLAM/multiprocess/mapper.py:212-250)
with the reference's UNMODIFIED `slerp` (SLAM/utils.py:593-652) on CPU tensors. Run in the build container, where
/root/reference exists:

    python tests/golden/make_history_merge_golden.py

`Mapping` itself cannot be constructed here (its constructor needs the dataset, the renderer and CUDA), so the eleven
assignments of `history_merge` are evaluated below on plain tensors exactly as written there -- including
`history_weight[0]` for the features and the scaling -- with `self.pointcloud.get_rotation` = F.normalize(_rotation)
(SLAM/gaussian_pointcloud.py:23,522-523). Inputs come from tests/helpers.py::history_merge_inputs (seeded); only the
reference's outputs are stored. Import stubs as in make_icp_golden.py.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_icp_golden import import_reference_icp  # noqa: E402


def reference_history_merge(slerp, hist, cur, max_weight):
    history_stat = {k: torch.from_numpy(v) for k, v in hist.items()}
    conf = torch.from_numpy(cur["confidence"])
    _xyz, _features_dc, _features_rest, _scaling = (torch.from_numpy(cur[k]) for k in ("xyz", "features_dc", "features_rest", "scaling"))
    get_rotation = torch.nn.functional.normalize(torch.from_numpy(cur["rotation_raw"]))
    history_weight = max_weight * history_stat["confidence"] / (conf + 1e-6)                       # mapper.py:215-219
    xyz_merge = history_stat["xyz"] * history_weight + (1 - history_weight) * _xyz                 # :223-226
    features_dc_merge = history_stat["features_dc"] * history_weight[0] + (1 - history_weight[0]) * _features_dc      # :228-231
    features_rest_merge = history_stat["features_rest"] * history_weight[0] + (1 - history_weight[0]) * _features_rest  # :233-236
    scaling_merge = history_stat["scaling"] * history_weight[0] + (1 - history_weight[0]) * _scaling                   # :238-241
    rotation_merge = slerp(history_stat["rotation"], get_rotation, 1 - history_weight)             # :242-244
    return {"xyz": xyz_merge, "features_dc": features_dc_merge, "features_rest": features_rest_merge, "scaling": scaling_merge,
            "rotation": rotation_merge}


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import helpers
    _, rutils = import_reference_icp()
    out = {}
    for name in helpers.HISTORY_MERGE_SIZES:
        hist, cur = helpers.history_merge_inputs(name)
        for mw in (0.5, 0.9):
            res = reference_history_merge(rutils.slerp, hist, cur, mw)
            for k, v in res.items():
                out[f"{name}_{mw}_{k}"] = v.numpy()
        out[f"{name}_checksum"] = np.array([float(np.sum(hist[k], dtype=np.float64)) for k in sorted(hist)]
                                           + [float(np.sum(cur[k], dtype=np.float64)) for k in sorted(cur)])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "history_merge.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
