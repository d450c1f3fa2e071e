"""Generates tests/golden/ssim.npz by running the UNMODIFIED `ssim` of the reference (utils/loss_utils.py:50-100, with its
`gaussian` / `create_window` / `_ssim`) on CPU tensors, as Mapping.loss_update calls it (mapper.py:411-415:
`1 - ssim(image.permute(2,0,1), gt.permute(2,0,1))`, i.e. un-batched (3,H,W) inputs). Run in the build container, where
/root/reference exists:

    python tests/golden/make_ssim_golden.py

Stored per case: the value 1 - ssim in float32 and float64 and the float64 gradient w.r.t. the first image (autograd
through the reference's own function). Inputs come from tests/helpers.py::ssim_inputs (seeded)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def reference_loss_utils():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)   # imports torch only
    return m


def main():
    import helpers
    lu = reference_loss_utils()
    out = {}
    for name in helpers.SSIM_CASES:
        a_np, b_np = helpers.ssim_inputs(name)
        a32, b32 = torch.from_numpy(a_np), torch.from_numpy(b_np)
        out[f"{name}_loss32"] = np.float32(1 - lu.ssim(a32, b32))
        a64 = a32.double().requires_grad_(True)
        l64 = 1 - lu.ssim(a64, b32.double())
        l64.backward()
        out[f"{name}_loss64"] = np.float64(l64.detach())
        out[f"{name}_grad64"] = a64.grad.numpy()
        out[f"{name}_checksum"] = np.array([a_np.sum(dtype=np.float64), b_np.sum(dtype=np.float64)])
    path = os.path.join(HERE, "ssim.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
