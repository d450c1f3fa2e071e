"""CPU tests of the SSIM chain: tests/golden/ssim.npz holds the outputs of the reference's UNMODIFIED `ssim`
(utils/loss_utils.py:50-100; tests/golden/make_ssim_golden.py). Held to it here:
* oracle/ssim_oracle.py::ssim_tiled -- the tile-for-tile restatement of csrc/ssim.cu (forward, derivative maps, backward);
* rtg_slam_b200.loss._ssim_term -- the torch restatement the GPU tests compare the kernels with (in float64).
So kernel == _ssim_term (GPU suite) and _ssim_term == reference ssim (here).

Tolerances: the reference rounds its 2-D window g_i * g_j to float32 (create_window), the separable form uses g_i and g_j
themselves -- a relative 6e-8 per weight. On textured images that moves the value by ~1e-7 and the gradient by ~1e-6 of
its largest entry; on smooth images sigma^2 = E[x^2] - mu^2 cancels next to C2 = 9e-4 and the same perturbation is
amplified ~200x ('smooth': 2e-4). Two float64 evaluations that differ only in that rounding are that far apart, so this is
the resolution at which "the same SSIM" is defined."""
GRAD_TOL = {"smooth": 1e-3}   # relative to the largest gradient entry; default 5e-6

import os

import numpy as np
import pytest
import torch

import helpers
from oracle.ssim_oracle import ssim_tiled

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ssim.npz"))


@pytest.mark.parametrize("name", sorted(helpers.SSIM_CASES))
def test_tiled_restatement_matches_reference_golden(name):
    a, b = helpers.ssim_inputs(name)
    assert np.allclose(GOLD[f"{name}_checksum"], [a.sum(dtype=np.float64), b.sum(dtype=np.float64)], rtol=1e-12)   # same inputs
    loss, grad = ssim_tiled(a.astype(np.float64), b.astype(np.float64))
    g = GOLD[f"{name}_grad64"]
    assert abs(loss - float(GOLD[f"{name}_loss64"])) < 5e-6
    assert np.abs(grad - g).max() < GRAD_TOL.get(name, 5e-6) * np.abs(g).max()
    assert abs(loss - float(GOLD[f"{name}_loss32"])) < 1e-5        # and the reference's own float32 evaluation is that close


@pytest.mark.parametrize("name", sorted(helpers.SSIM_CASES))
def test_torch_restatement_used_by_the_gpu_tests_matches_reference_golden(name):
    from rtg_slam_b200.loss import _ssim_term
    a_np, b_np = helpers.ssim_inputs(name)
    a = torch.from_numpy(a_np).double().requires_grad_(True)
    loss = _ssim_term(a, torch.from_numpy(b_np).double())
    loss.backward()
    g = GOLD[f"{name}_grad64"]
    assert abs(float(loss.detach()) - float(GOLD[f"{name}_loss64"])) < 5e-6      # float64 window here, float32-rounded in the reference
    assert np.abs(a.grad.numpy() - g).max() < GRAD_TOL.get(name, 5e-6) * np.abs(g).max()
    # in float32 the restatement builds the very same window: same value
    assert abs(float(_ssim_term(torch.from_numpy(a_np), torch.from_numpy(b_np))) - float(GOLD[f"{name}_loss32"])) < 2e-6
