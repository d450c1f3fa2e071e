"""Diagnose full-size (1 M Gaussians, 1200x680) differences between the product path and the CPU oracle."""
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
mask = scene.random_tile_mask(cam, 0.5, seed=11) if (len(sys.argv) > 2 and sys.argv[2] == "mask") else None
grads = scene.upstream_grads(cam, seed=5)
r = helpers.run_ours(cam, g, dev, tile_mask=mask, grads=grads)
o = OracleRender(cam, g, tile_mask=mask, precision="f32", tie_eps=1e-4, nthreads=32)
names = ("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii")
od = dict(zip(names, o.outputs()))
print("R oracle", o.num_rendered)
try:
    print(helpers.compare_outputs(r, od, tie=o.tie, max_bad_frac=1.0))
except Exception as e:
    print("compare failed", e)
bad = (r["hit_depth"][0] != od["hit_depth"][0]) | (np.abs(r["depth"][0] - od["depth"][0]) > 1e-4) | (np.abs(r["color"] - od["color"]).max(0) > 1e-4)
bad &= ~o.tie.astype(bool)
ys, xs = np.nonzero(bad)
print("bad pixels", bad.sum(), "in tiles", len(set(zip(ys // 16, xs // 16))))
pl, rg = o.binning()
lens = rg[:, 1] - rg[:, 0]
print("oracle tile list: max", lens.max(), "n>4096:", (lens > 4096).sum())
tiles = sorted(set(zip(ys // 16, xs // 16)))[:10]
tw = cam.tile_grid[1]
for ty, tx in tiles:
    t = ty * tw + tx
    m = bad[ty*16:(ty+1)*16, tx*16:(tx+1)*16]
    print("tile", ty, tx, "len", lens[t], "bad px", m.sum())
for y, x in list(zip(ys, xs))[:8]:
    print((y, x), "ours hit", r["hit_depth"][0, y, x], "oracle", od["hit_depth"][0, y, x], "depth", r["depth"][0, y, x], od["depth"][0, y, x],
          "T", r["T_map"][0, y, x], od["T_map"][0, y, x], "col", r["color"][:, y, x], od["color"][:, y, x])
og = o.backward(*grads, nthreads=16)
for k in ("means3D", "shs", "opacities", "scales", "rotations"):
    print("grad", k, helpers.rel_err(r["grads"][k], og[k]))
if mask is not None:
    up = np.kron(mask, np.ones((16, 16), np.int32))[: cam.height, : cam.width].astype(bool)
else:
    up = np.ones((cam.height, cam.width), bool)
for tag, dd in (("ours", r), ("oracle", od)):
    hit = dd["hit_depth"][0]; d = dd["depth"][0]
    v1 = (hit >= 0) & up & ~(d > 0); v2 = (d > 0) & ~((hit >= 0) & up)
    print(tag, "hit&depth<=0:", v1.sum(), " depth>0&nohit:", v2.sum(), " frac ok", ((d > 0) == ((hit >= 0) & up)).mean())
    ys, xs = np.nonzero(v1 | v2)
    for y, x in list(zip(ys, xs))[:5]:
        print("   ", (y, x), "hit", hit[y, x], "depth", d[y, x], "up", up[y, x], "T", dd["T_map"][0, y, x])
