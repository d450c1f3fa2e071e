"""Generates tests/golden/mapstats_tilemasks.npz by running the UNMODIFIED reference tile-mask builders
(SLAM/utils.py:681-734: pixelmask2tilemask, transmission2tilemask, colorerror2tilemask) on CPU tensors. Run in the
build container, where /root/reference exists:

    python tests/golden/make_mapstats_golden.py

Import stubs as in make_icp_golden.py. Sizes: 680x1200 (Replica, last tile row half empty) and 77x45 (ragged);
inputs come from tests/helpers.py::mapstats_inputs (seeded), only the reference's outputs are stored.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_icp_golden import import_reference_icp  # noqa: E402


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import helpers
    _, rutils = import_reference_icp()
    out = {}
    for name in helpers.MAPSTATS_SIZES:
        T, err = helpers.mapstats_inputs(name)   # inputs are regenerated from the seed by the tests: only outputs are stored
        pixelmask = torch.from_numpy(T != 1)
        out[f"{name}_pix2tile"] = rutils.pixelmask2tilemask(pixelmask, 16).numpy()
        for ratio in (0.5, 0.1):
            out[f"{name}_trans_{ratio}"] = rutils.transmission2tilemask(pixelmask, 16, ratio).numpy()
        for ratio in (0.4, 0.05):
            out[f"{name}_cerr_{ratio}"] = rutils.colorerror2tilemask(torch.from_numpy(err), 16, ratio).numpy()
        out[f"{name}_checksum"] = np.array([float(T.sum(dtype=np.float64)), float(err.sum(dtype=np.float64))])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mapstats_tilemasks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
