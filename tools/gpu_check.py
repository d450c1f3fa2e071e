"""First-light diagnostics on the GPU box: product path vs CPU oracle vs the reference's own CUDA build.
Prints one block per scene; never stops at the first failure."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from rtg_slam_b200 import scene  # noqa: E402
from oracle.splat_oracle import OracleRender  # noqa: E402

dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), "ref ext:", helpers.ref_cuda_module() is not None, flush=True)


def stats(tag, a, b, tie=None):
    try:
        s = helpers.compare_outputs(a, b, tie=tie, label=tag, max_bad_frac=1.0)
        print(f"  {tag}: " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in s.items()), flush=True)
    except Exception:
        traceback.print_exc()


def gstats(tag, a, b):
    print(f"  {tag}: " + " ".join(f"{k}={helpers.rel_err(a[k], b[k]):.2e}" for k in ("means3D", "shs", "opacities", "scales", "rotations")),
          flush=True)


def raster_case(name, cam, g, mask=None, use_oracle=True):
    print(f"== {name}: P={g['xyz'].shape[0]} {cam.width}x{cam.height} mask={'yes' if mask is not None else 'no'}", flush=True)
    grads = scene.upstream_grads(cam, seed=5)
    try:
        ours = helpers.run_ours(cam, g, dev, tile_mask=mask, grads=grads)
    except Exception:
        traceback.print_exc()
        return
    ref = None
    if helpers.ref_cuda_module() is not None:
        try:
            ref = helpers.run_ref_cuda(cam, g, dev, tile_mask=mask, grads=grads)
            ref2 = helpers.run_ref_cuda(cam, g, dev, tile_mask=mask, grads=grads)
            print(f"  ref: R={ref['num_rendered']} tiles={ref['num_tile']} vis={(ref['radii'] > 0).sum()}", flush=True)
            stats("ours-vs-ref", ours, ref)
            gstats("grads ours-vs-ref", ours["grads"], ref["grads"])
            gstats("grads ref-vs-ref (jitter)", ref2["grads"], ref["grads"])
        except Exception:
            traceback.print_exc()
    if use_oracle:
        try:
            o = OracleRender(cam, g, tile_mask=mask, precision="f32", tie_eps=1e-4)
            od = dict(zip(("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii"), o.outputs()))
            stats("ours-vs-oracle", ours, od, tie=o.tie)
            og = o.backward(*grads)
            gstats("grads ours-vs-oracle", ours["grads"], og)
            if ref is not None:
                stats("ref-vs-oracle", ref, od, tie=o.tie)
                gstats("grads ref-vs-oracle", ref["grads"], og)
        except Exception:
            traceback.print_exc()


def timing(P, camname, iters=10):
    cam = scene.make_camera(camname)
    g = scene.surfel_room(P, seed=2024)
    from rtg_slam_b200.rasterizer import GaussianRasterizer
    rs = helpers.make_settings(cam, dev)
    t = helpers.to_torch(g, dev)
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("xyz", "shs", "opacity", "scales", "rotations")}
    gc, gd = [torch.from_numpy(x).to(dev) for x in scene.upstream_grads(cam)]
    rast = GaussianRasterizer(rs)

    def step():
        out = rast(means3D=leaves["xyz"], opacities=leaves["opacity"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward([out[0], out[1]], [gc, gd])
        return out
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    vis = int((out[7] > 0).sum())
    print(f"== timing ours P={P} {camname}: {ms:.3f} ms fwd+bwd, vis={vis}", flush=True)
    mod = helpers.ref_cuda_module()
    if mod is not None:
        try:
            grads = scene.upstream_grads(cam)
            for _ in range(2):
                helpers.run_ref_cuda(cam, g, dev, grads=grads)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(3):
                r = helpers.run_ref_cuda(cam, g, dev, grads=grads)
            torch.cuda.synchronize()
            print(f"   ref (incl. host copies of outputs): {(time.time() - t0) / 3 * 1e3:.2f} ms, R={r['num_rendered']}", flush=True)
        except Exception:
            traceback.print_exc()


def icp_check():
    from oracle import icp_oracle as io
    from rtg_slam_b200 import icp as ricp
    for name in ("icp_small", "icp_ragged"):
        try:
            gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
            K = tuple(float(k) for k in gold["K"])
            d0 = torch.from_numpy(gold["depth0"]).to(dev)
            d1 = torch.from_numpy(gold["depth1"]).to(dev)
            v0, n0 = ricp.build_pyramids(d0, K, 3)
            v1, n1 = ricp.build_pyramids(d1, K, 3)
            for i in range(3):
                print(f"  {name} pyr{i}: v {np.abs(v0[i].cpu().numpy() - gold[f'v0_{i}']).max():.2e} n {np.abs(n0[i].cpu().numpy() - gold[f'n0_{i}']).max():.2e}")
            pose = torch.eye(4, device=dev)
            for lvl, s in enumerate((0.25, 0.5, 1.0)):
                tr = ricp.ICP(5, damping=1e-4, distance_threshold=0.1, normal_threshold=20)
                Kl = tuple(float(np.float32(k) * np.float32(s)) for k in K)
                pose, vr = tr.icp(pose, v1[lvl], v0[lvl], n1[lvl], n0[lvl], Kl)
                print(f"  {name} level{lvl}: |dT|_F vs golden {np.linalg.norm(pose.cpu().numpy() - gold['poses'][lvl]):.2e} valid {float(vr):.4f}/{gold['valid_ratios'][lvl]:.4f}")
        except Exception:
            traceback.print_exc()


def adam_check():
    from rtg_slam_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(1000, 3), (1000, 1, 3), (1000, 15, 3), (1000, 1), (1000, 3), (1000, 4), (7,)]
    lrs = [1e-3, 5e-4, 2.5e-5, 0.0, 4e-3, 1e-3, 1e-2]
    pa = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    ob = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for it in range(10):
        gs = [torch.randn(s, device=dev) * (10 ** np.random.uniform(-4, 0)) for s in shapes]
        for p, q, g in zip(pa, pb, gs):
            p.grad = g.clone(); q.grad = g.clone()
        oa.step(); ob.step()
    print("== adam: max rel err after 10 steps", max(float((p - q).abs().max() / q.abs().max()) for p, q in zip(pa, pb)), flush=True)


if __name__ == "__main__":
    cam = scene.make_camera("small")
    raster_case("room-small", cam, scene.surfel_room(3000, seed=1))
    raster_case("room-masked", scene.make_camera("small", c2w=scene.small_pose()), scene.surfel_room(3000, seed=2),
                mask=scene.random_tile_mask(scene.make_camera("small"), 0.5, seed=102))
    raster_case("blobs-ragged", scene.make_camera("ragged"), scene.random_blobs(1500, seed=7))
    raster_case("blobs-tiny", scene.make_camera("tiny", c2w=scene.small_pose()), scene.random_blobs(300, seed=9))
    raster_case("room-tum-10k", scene.make_camera("tum"), scene.surfel_room(10000, seed=2024))
    raster_case("room-replica-100k", scene.make_camera("replica"), scene.surfel_room(100000, seed=2024))
    try:
        icp_check()
    except Exception:
        traceback.print_exc()
    try:
        adam_check()
    except Exception:
        traceback.print_exc()
    for P, c in ((300000, "replica"), (1000000, "replica")):
        try:
            timing(P, c)
        except Exception:
            traceback.print_exc()
