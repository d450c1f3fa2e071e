"""Generates tests/golden/frameprep_*.npz by running the UNMODIFIED reference functions of SLAM/utils.py
(bilateralFilter_torch, compute_vertex_map, compute_normal_map, compute_confidence_map) in the order of
Tracker.map_preprocess (SLAM/multiprocess/tracker.py:104-132) on CPU tensors. Run in the build container:

    python tests/golden/make_frameprep_golden.py

Import stubs as in make_icp_golden.py; compute_confidence_map calls `.cuda()` on its grids (utils.py:132-134), which is
made a no-op for the duration of the call -- nothing else of the reference is touched. Inputs come from
tests/helpers.py::frameprep_inputs (seeded); only the reference's outputs are stored (fp16-exact inputs are not needed).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_icp_golden import import_reference_icp  # noqa: E402


def reference_map_preprocess(rutils, depth, K, depth_filter, min_depth, max_depth, thresh):
    depth_map = torch.from_numpy(depth)[..., None].clone()
    intrinsic = torch.tensor(K, dtype=torch.float32)
    depth_map_filter = rutils.bilateralFilter_torch(depth_map, 5, 2, 2) if depth_filter else depth_map   # tracker.py:107-110
    valid = (depth_map_filter > min_depth) & (depth_map_filter < max_depth)                               # :112
    depth_map_filter[~valid] = 0.0
    vertex_map_c = rutils.compute_vertex_map(depth_map_filter, intrinsic)                                 # :117
    normal_map_c = rutils.compute_normal_map(vertex_map_c)
    cuda, torch.Tensor.cuda = torch.Tensor.cuda, lambda self, *a, **k: self
    try:
        confidence_map = rutils.compute_confidence_map(normal_map_c, intrinsic)
    finally:
        torch.Tensor.cuda = cuda
    bad = ((normal_map_c == 0).all(dim=-1)) | (confidence_map < thresh)[..., 0]                           # :122-124
    depth_map_filter[bad] = 0
    normal_map_c[bad] = 0
    vertex_map_c[bad] = 0
    confidence_map[bad] = 0
    return dict(depth_map=depth_map_filter.numpy(), vertex_map_c=vertex_map_c.numpy(), normal_map_c=normal_map_c.numpy(),
                confidence_map=confidence_map.numpy(), invalid_confidence_mask=bad.numpy())


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import helpers
    _, rutils = import_reference_icp()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, cfg in helpers.FRAMEPREP_CASES.items():
        depth, K = helpers.frameprep_inputs(name)
        out = reference_map_preprocess(rutils, depth, K, cfg["depth_filter"], cfg["min_depth"], cfg["max_depth"], cfg["thresh"])
        out["checksum"] = np.array([float(depth.sum(dtype=np.float64))])
        # fp16 would lose the parity bar; keep fp32 but only every stored array is small (<= 120x160)
        np.savez_compressed(os.path.join(here, f"frameprep_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()}, "invalid fraction", float(out["invalid_confidence_mask"].mean()))


if __name__ == "__main__":
    main()
