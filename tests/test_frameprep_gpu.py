"""GPU parity of the tracker-side frame preprocessing through the C ABI: against the golden outputs of
the unmodified reference functions and against the numpy oracle."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import frameprep_oracle as fo
from rtg_slam_b200 import frameprep, scene

pytestmark = pytest.mark.gpu
TOL = {"depth_map": 3e-6, "vertex_map_c": 6e-6, "normal_map_c": 2e-4, "confidence_map": 2e-4}


def _compare(out, want, thresh):
    bad, wbad = out["invalid_confidence_mask"].cpu().numpy(), want["invalid_confidence_mask"]
    diff = bad != wbad
    assert diff.mean() <= 1e-3
    if diff.any():  # only pixels whose confidence sits at the threshold may be classified differently
        c = np.where(bad, want["confidence_map"][..., 0], out["confidence_map"][..., 0].cpu().numpy())  # the unmasked side's value
        assert np.all(np.abs(c[diff] - thresh) < 1e-3)
    same = ~diff
    for k, tol in TOL.items():
        o, w = out[k].cpu().numpy(), want[k]
        assert o.shape == w.shape, k
        assert np.abs(o - w)[same].max() <= tol, k


@pytest.mark.parametrize("name", sorted(helpers.FRAMEPREP_CASES))
def test_map_preprocess_matches_reference_golden_and_oracle(cuda_device, name):
    cfg = helpers.FRAMEPREP_CASES[name]
    depth, K = helpers.frameprep_inputs(name)
    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", f"frameprep_{name}.npz")))
    out = frameprep.map_preprocess_maps(torch.from_numpy(depth).to(cuda_device)[..., None], K, cfg["depth_filter"], cfg["min_depth"],
                                        cfg["max_depth"], cfg["thresh"])
    assert out["depth_map"].shape == (depth.shape[0], depth.shape[1], 1) and out["invalid_confidence_mask"].dtype == torch.bool
    _compare(out, gold, cfg["thresh"])
    _compare(out, fo.map_preprocess(depth, K, cfg["depth_filter"], cfg["min_depth"], cfg["max_depth"], cfg["thresh"]), cfg["thresh"])


def test_full_size_frame_and_standalone_filter(cuda_device):
    cam = scene.make_camera("replica")
    rng = np.random.default_rng(5)
    depth = scene.raycast_room_depth(cam).astype(np.float32) + rng.normal(0, 0.002, (cam.height, cam.width)).astype(np.float32)
    depth[rng.uniform(size=depth.shape) < 0.02] = 0
    K = [[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]]
    d = torch.from_numpy(depth).to(cuda_device)
    out = frameprep.map_preprocess_maps(d, K, True, 0.3, 5.0, 0.2)
    bad = out["invalid_confidence_mask"]
    # size-independent properties: masked pixels are zero everywhere, kept normals are unit, confidence in [thresh, 1]
    assert float(out["depth_map"][..., 0][bad].abs().max()) == 0 and float(out["normal_map_c"][bad].abs().max()) == 0
    assert float(out["vertex_map_c"][bad].abs().max()) == 0 and float(out["confidence_map"][..., 0][bad].abs().max()) == 0
    nn = out["normal_map_c"][~bad].norm(dim=-1)   # n / (|n| + 1e-8): unit unless the Sobel cross product is tiny
    assert float(nn.max()) <= 1 + 1e-5 and float((nn > 0.999).float().mean()) > 0.99 and float(nn.min()) > 0
    c = out["confidence_map"][..., 0][~bad]
    assert float(c.min()) >= 0.2 and float(c.max()) <= 1 + 1e-6
    assert torch.allclose(out["vertex_map_c"][..., 2], out["depth_map"][..., 0])
    assert 0.3 < float((~bad).float().mean()) <= 1.0
    # the filter alone, against the oracle on a crop (the oracle is slow at full size)
    f = frameprep.bilateralFilter_torch(d[:96, :128].contiguous(), 5, 2, 2)[..., 0].cpu().numpy()
    assert np.abs(f - fo.bilateral_filter(depth[:96, :128], 5, 2, 2)).max() < 3e-6
    with pytest.raises(frameprep._lib.RtgError):
        frameprep.bilateralFilter_torch(d, 99, 2, 2)
