"""Generates tests/golden/history_merge.npz by EXECUTING the reference's own `Mapping.history_merge`
(SLAM/multiprocess/mapper.py:212-250) with the reference's unmodified `slerp` (SLAM/utils.py:593-652) on CPU tensors. Run in
the build container, where /root/reference exists:

    python tests/golden/make_history_merge_golden.py

The module SLAM/multiprocess/mapper.py cannot be imported here (tensorboard, the CUDA extensions, the dataset classes) and
`Mapping` cannot be constructed without a dataset, so the method is taken from the file as it is: its `def` is located
with `ast`, compiled unchanged and called with a duck-typed `self` that offers the five attributes it reads
(`pointcloud.get_confidence / get_xyz / get_rotation / _features_dc / _features_rest / _scaling`, `verbose`), with
`get_rotation = F.normalize(_rotation)` as SLAM/gaussian_pointcloud.py:23,522-523 defines it. No line of the method is restated
in this repository. Inputs come from tests/helpers.py::history_merge_inputs (seeded); only the outputs are stored.
Import stubs for `slerp`'s module as in make_icp_golden.py.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_icp_golden import import_reference_icp  # noqa: E402

MAPPER = "/root/reference/SLAM/multiprocess/mapper.py"


def reference_method(name, namespace):
    """The function object of `Mapping.<name>`, compiled from the reference file's own source, unmodified."""
    tree = ast.parse(open(MAPPER).read(), MAPPER)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Mapping")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    exec(compile(mod, MAPPER, "exec"), namespace)   # line numbers of the traceback stay those of the reference file
    return namespace[name]


def reference_history_merge(history_merge, hist, cur, max_weight):
    pc = types.SimpleNamespace(
        get_confidence=torch.from_numpy(cur["confidence"]), get_xyz=torch.from_numpy(cur["xyz"]),
        get_rotation=torch.nn.functional.normalize(torch.from_numpy(cur["rotation_raw"])),
        _xyz=torch.from_numpy(cur["xyz"]), _features_dc=torch.from_numpy(cur["features_dc"]),
        _features_rest=torch.from_numpy(cur["features_rest"]), _scaling=torch.from_numpy(cur["scaling"]),
        _rotation=torch.from_numpy(cur["rotation_raw"]))
    fake_self = types.SimpleNamespace(pointcloud=pc, verbose=False)
    history_merge(fake_self, {k: torch.from_numpy(v) for k, v in hist.items()}, max_weight)
    return {"xyz": pc._xyz, "features_dc": pc._features_dc, "features_rest": pc._features_rest, "scaling": pc._scaling,
            "rotation": pc._rotation}


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import helpers
    _, rutils = import_reference_icp()
    history_merge = reference_method("history_merge", {"torch": torch, "slerp": rutils.slerp})
    out = {}
    for name in helpers.HISTORY_MERGE_SIZES:
        hist, cur = helpers.history_merge_inputs(name)
        for mw in (0.5, 0.9):
            res = reference_history_merge(history_merge, hist, cur, mw)
            for k, v in res.items():
                out[f"{name}_{mw}_{k}"] = v.numpy()
        out[f"{name}_checksum"] = np.array([float(np.sum(hist[k], dtype=np.float64)) for k in sorted(hist)]
                                           + [float(np.sum(cur[k], dtype=np.float64)) for k in sorted(cur)])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "history_merge.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
