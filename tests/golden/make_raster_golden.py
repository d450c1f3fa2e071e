"""Generates tests/golden/raster_*.npz from the reference's OWN CUDA rasterizer (oracle/_ref/_C_depth*.so, the
unmodified RAST sources compiled for sm_100 by oracle/build_ref.py). Needs a GPU:

    gpurun -- 'python tests/golden/make_raster_golden.py gpurun_out/golden'   # then copy the .npz files here

Each file holds the scene recipe (name, P, seed, camera, mask seed), the 8 forward outputs, and the gradients
for the fixed upstream gradients of rtg_slam_b200.scene.upstream_grads. Inputs are regenerated from the recipe
by the tests (numpy Generator streams are stable across versions), so the fixtures stay small.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (scene fn, P, seed, camera, pose, tile-mask keep fraction or None)
    "raster_room_small": ("surfel_room", 3000, 1, "small", None, None),
    "raster_room_masked": ("surfel_room", 3000, 2, "small", "small_pose", 0.5),
    "raster_blobs_ragged": ("random_blobs", 1500, 7, "ragged", None, None),
    "raster_blobs_tiny": ("random_blobs", 300, 9, "tiny", "small_pose", None),
}


def build_case(name):
    from rtg_slam_b200 import scene
    fn, P, seed, camname, pose, keep = CASES[name]
    cam = scene.make_camera(camname, c2w=scene.small_pose() if pose else None)
    g = getattr(scene, fn)(P, seed=seed)
    mask = scene.random_tile_mask(cam, keep, seed=seed + 100) if keep else None
    grads = scene.upstream_grads(cam, seed=seed + 200)
    return cam, g, mask, grads


def main():
    import helpers
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda", 0)
    for name in CASES:
        cam, g, mask, grads = build_case(name)
        r = helpers.run_ref_cuda(cam, g, dev, tile_mask=mask, grads=grads)
        r2 = helpers.run_ref_cuda(cam, g, dev, tile_mask=mask, grads=grads)  # atomics jitter of the reference itself
        jit = {k: helpers.rel_err(r["grads"][k], r2["grads"][k]) for k in r["grads"]}
        np.savez_compressed(
            os.path.join(out, name + ".npz"),
            **{k: r[k] for k in ("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii")},
            num_rendered=np.int64(r["num_rendered"]), num_tile=np.int64(r["num_tile"]),
            **{"grad_" + k: v for k, v in r["grads"].items()}, **{"jitter_" + k: np.float64(v) for k, v in jit.items()})
        print(name, "R", r["num_rendered"], "tiles", r["num_tile"], "jitter", {k: f"{v:.1e}" for k, v in jit.items()})


if __name__ == "__main__":
    main()
