"""Run the forward/backward after poisoning the caching allocator's free blocks, compare with the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from rtg_slam_b200 import scene
from oracle.splat_oracle import OracleRender
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
camname = sys.argv[2] if len(sys.argv) > 2 else "replica"
poison = len(sys.argv) <= 3 or sys.argv[3] != "nopoison"
cam = scene.make_camera(camname)
g = scene.surfel_room(P, seed=2024)
mask = scene.random_tile_mask(cam, 0.5, seed=11)
grads = scene.upstream_grads(cam, seed=5)
if poison:
    junk = [torch.full((n,), 0x7f7f7f7f, dtype=torch.int32, device=dev) for n in (1 << 28, 1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14, 4096, 1024, 256, 64)]
    junk2 = [torch.full((n,), 0x7f7f7f7f, dtype=torch.int32, device=dev) for n in (1 << 25, 1 << 23, 1 << 21, 1 << 19, 1 << 17, 1 << 15, 1 << 13)]
    del junk, junk2
r = helpers.run_ours(cam, g, dev, tile_mask=mask, grads=grads)
o = OracleRender(cam, g, tile_mask=mask, precision="f32", tie_eps=1e-4, nthreads=32)
names = ("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii")
od = dict(zip(names, o.outputs()))
try:
    print(helpers.compare_outputs(r, od, tie=o.tie, max_bad_frac=1.0))
except Exception as e:
    print("compare failed", e)
og = o.backward(*grads, nthreads=16)
for k in ("means3D", "shs", "opacities", "scales", "rotations"):
    print("grad", k, helpers.rel_err(r["grads"][k], og[k]), "finite", np.isfinite(r["grads"][k]).all())
