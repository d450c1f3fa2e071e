"""Where do the product path and the reference CUDA rasterizer differ at the benchmark size? (forward only)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from rtg_slam_b200 import scene
from oracle.splat_oracle import OracleRender
dev = torch.device("cuda", 0)
cam = scene.make_camera("replica")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = scene.surfel_room(P, seed=2024)
ours = helpers.run_ours(cam, g, dev)
ref = helpers.run_ref_cuda(cam, g, dev)
dT = np.abs(ours["T_map"][0].astype(np.float64) - ref["T_map"][0])
dC = np.abs(ours["color"].astype(np.float64) - ref["color"]).max(0)
same = (ours["hit_color"][0] == ref["hit_color"][0]) & (ours["hit_depth"][0] == ref["hit_depth"][0])
print("index mismatches", (~same).sum(), "max dT", dT.max(), "max dC", dC.max())
for thr in (1e-6, 1e-5, 1e-4, 1e-3):
    print(f"pixels with dT > {thr}: {(dT > thr).sum()}   dC > {thr}: {(dC > thr).sum()}")
o = OracleRender(cam, g, precision="f32", tie_eps=1e-4, nthreads=64)
od = dict(zip(("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii"), o.outputs()))
tie = o.tie.astype(bool)
bad = dT > 1e-4
print("bad pixels", bad.sum(), "of which oracle-flagged ties", (bad & tie).sum())
print("oracle vs ref: max dT", np.abs(od["T_map"][0] - ref["T_map"][0]).max(), " oracle vs ours:", np.abs(od["T_map"][0] - ours["T_map"][0]).max())
ys, xs = np.nonzero(bad)
for y, x in list(zip(ys, xs))[:12]:
    print((y, x), "T ours/ref/oracle", ours["T_map"][0, y, x], ref["T_map"][0, y, x], od["T_map"][0, y, x], "ratio ours/ref", ours["T_map"][0, y, x] / ref["T_map"][0, y, x],
          "tie", tie[y, x], "hit", ours["hit_depth"][0, y, x], "hcw", ours["hit_color_weight"][0, y, x], ref["hit_color_weight"][0, y, x])
