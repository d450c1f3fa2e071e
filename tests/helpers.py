"""Shared test helpers: scene -> torch tensors, calls into the product path, the oracle and (when its
extension has been built into oracle/_ref/) the reference's own CUDA rasterizer."""
from __future__ import annotations

import glob
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFAULT_SETTINGS = dict(scale_modifier=1.0, color_sigma=3.0, opaque_threshold=0.6, depth_threshold=1.0,
                        normal_threshold=float(np.cos(np.deg2rad(60.0))), T_threshold=1e-4, sh_degree=3)


def to_torch(g, device):
    return {k: torch.from_numpy(v).to(device) for k, v in g.items()}


def make_settings(cam, device, **over):
    from rtg_slam_b200.rasterizer import GaussianRasterizationSettings
    st = dict(DEFAULT_SETTINGS)
    st.update(over)
    return GaussianRasterizationSettings(
        image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.tensor(st.get("bg", (0.0, 0.0, 0.0)), dtype=torch.float32, device=device),
        scale_modifier=st["scale_modifier"],
        viewmatrix=torch.from_numpy(cam.viewmatrix).to(device), projmatrix=torch.from_numpy(cam.projmatrix).to(device),
        sh_degree=st["sh_degree"], campos=torch.from_numpy(cam.campos).to(device),
        opaque_threshold=st["opaque_threshold"], normal_threshold=st["normal_threshold"], depth_threshold=st["depth_threshold"],
        prefiltered=False, debug=False, cx=cam.cx, cy=cam.cy, color_sigma=st["color_sigma"], T_threshold=st["T_threshold"])


def run_ours(cam, g, device, tile_mask=None, grads=None, **over):
    """Forward (and backward if `grads`=(dL_dcolor, dL_ddepth) numpy) through the public operator API."""
    from rtg_slam_b200.rasterizer import GaussianRasterizer
    rs = make_settings(cam, device, **over)
    t = to_torch(g, device)
    leaves = {k: t[k].clone().requires_grad_(grads is not None) for k in ("xyz", "shs", "opacity", "scales", "rotations")}
    tm = None if tile_mask is None else torch.from_numpy(np.ascontiguousarray(tile_mask, dtype=np.int32)).to(device)
    out = GaussianRasterizer(rs)(means3D=leaves["xyz"], opacities=leaves["opacity"], shs=leaves["shs"], scales=leaves["scales"],
                                 rotations=leaves["rotations"], tile_mask=tm)
    res = dict(zip(("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map", "radii"),
                   [o.detach().cpu().numpy() for o in out]))
    if grads is not None:
        gc = torch.from_numpy(grads[0]).to(device)
        gd = torch.from_numpy(grads[1]).to(device)
        loss = (out[0] * gc).sum() + (out[1] * gd).sum()
        loss.backward()
        res["grads"] = dict(means3D=leaves["xyz"].grad.cpu().numpy(), shs=leaves["shs"].grad.cpu().numpy(),
                            opacities=leaves["opacity"].grad.cpu().numpy(), scales=leaves["scales"].grad.cpu().numpy(),
                            rotations=leaves["rotations"].grad.cpu().numpy())
    return res


# ---------------------------------------------------------------- reference CUDA (oracle/_ref)
_REF = None


def ref_cuda_module():
    """The reference's own `_C_depth` extension, built unmodified by oracle/build_ref.py. None if absent."""
    global _REF
    if _REF is None:
        so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_C_depth*.so"))
        if not so:
            _REF = False
        else:
            spec = importlib.util.spec_from_file_location("_C_depth", so[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _REF = mod
    return _REF or None


def run_ref_cuda(cam, g, device, tile_mask=None, grads=None, **over):
    """Calls rasterize_gaussians / rasterize_gaussians_backward of the reference exactly as its python shim does
    (RAST/diff_gaussian_rasterization_depth/__init__.py:69-97,200-232)."""
    mod = ref_cuda_module()
    assert mod is not None
    st = dict(DEFAULT_SETTINGS)
    st.update(over)
    t = to_torch(g, device)
    H, W = cam.height, cam.width
    th, tw = (H + 15) // 16, (W + 15) // 16
    tm = torch.ones((th, tw), dtype=torch.int32, device=device) if tile_mask is None else \
        torch.from_numpy(np.ascontiguousarray(tile_mask, dtype=np.int32)).to(device)
    bg = torch.tensor(st.get("bg", (0.0, 0.0, 0.0)), dtype=torch.float32, device=device)
    vm = torch.from_numpy(cam.viewmatrix).to(device)
    pm = torch.from_numpy(cam.projmatrix).to(device)
    cp = torch.from_numpy(cam.campos).to(device)
    e = torch.Tensor([])
    args = (bg, t["xyz"], e, t["opacity"], t["scales"], t["rotations"], st["scale_modifier"], e, vm, pm, tm, cam.tanfovx, cam.tanfovy,
            H, W, cam.cx, cam.cy, t["shs"], st["sh_degree"], st["color_sigma"], cp, st["opaque_threshold"], st["depth_threshold"],
            st["normal_threshold"], st["T_threshold"], False, False)
    (num_rendered, num_tile, color, depth, hit_color, hit_depth, hcw, hdw, T_map, radii, geomB, binB, imgB, tile_indices) = \
        mod.rasterize_gaussians(*args)
    res = dict(color=color, depth=depth, hit_color=hit_color, hit_depth=hit_depth, hit_color_weight=hcw, hit_depth_weight=hdw,
               T_map=T_map, radii=radii)
    res = {k: v.cpu().numpy() for k, v in res.items()}
    res["num_rendered"], res["num_tile"] = num_rendered, num_tile
    if grads is not None:
        gc = torch.from_numpy(grads[0]).to(device)
        gd = torch.from_numpy(grads[1]).to(device)
        bargs = (tile_indices, num_tile, bg, t["xyz"], radii, e, t["scales"], t["rotations"], st["scale_modifier"], e, vm, pm,
                 cam.tanfovx, cam.tanfovy, cam.cx, cam.cy, st["depth_threshold"], st["normal_threshold"], gc, gd, t["shs"],
                 st["sh_degree"], cp, geomB, num_rendered, binB, imgB, hit_depth, False)
        (g2d, gcol, gop, gm3, gcov, gsh, gsc, grot) = mod.rasterize_gaussians_backward(*bargs)
        res["grads"] = dict(means3D=gm3.cpu().numpy(), shs=gsh.cpu().numpy(), opacities=gop.cpu().numpy(), scales=gsc.cpu().numpy(),
                            rotations=grot.cpu().numpy(), means2D=g2d.cpu().numpy(), colors=gcol.cpu().numpy(),
                            cov3D=gcov.cpu().numpy())
    return res


# ---------------------------------------------------------------- comparison
def rel_err(a, b):
    """max|a-b| / (max|b| + 1e-12): the per-tensor gradient metric of SURVEY.md section 8(c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12)) if a.size else 0.0


def compare_outputs(a, b, tie=None, tol=1e-4, max_bad_frac=2e-4, label=""):
    """Float maps: L_inf < tol on all pixels outside `tie` and on all but `max_bad_frac` of the pixels overall.
    Index maps: exact outside `tie`, same outlier allowance. Returns a dict of statistics; raises AssertionError."""
    H, W = a["color"].shape[-2:]
    ok = np.ones((H, W), bool) if tie is None else ~tie.astype(bool)
    idx_equal = (a["hit_color"][0] == b["hit_color"][0]) & (a["hit_depth"][0] == b["hit_depth"][0])
    stats = {}
    bad = ~idx_equal
    for k in ("color", "depth", "hit_color_weight", "hit_depth_weight", "T_map"):
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max(axis=0)
        stats[k] = float(d[ok & idx_equal].max()) if (ok & idx_equal).any() else 0.0
        bad |= d >= tol
    stats["bad_pixels"] = int((bad & ok).sum())
    stats["tie_pixels"] = int((~ok).sum())
    stats["radii_mismatch"] = int((a["radii"] != b["radii"]).sum())
    frac = stats["bad_pixels"] / float(H * W)
    assert frac <= max_bad_frac, f"{label}: {stats['bad_pixels']} pixels differ beyond tol ({frac:.2e} of the image): {stats}"
    return stats


# ---------------------------------------------------------------- map statistics
MAPSTATS_SIZES = {"replica": (680, 1200), "ragged": (77, 45)}


def mapstats_inputs(name):
    """Seeded inputs of the tile-mask goldens (tests/golden/make_mapstats_golden.py regenerates them the same way): a
    transmittance-like map (1 where nothing was rendered) and a colour-error image."""
    H, W = MAPSTATS_SIZES[name]
    rng = np.random.default_rng(2024 + H)
    yy, xx = np.mgrid[0:H, 0:W]
    T = np.ones((H, W), np.float32)
    for _ in range(12):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(5, 0.3 * min(H, W))
        T[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = rng.uniform(0, 0.5)
    T[rng.uniform(size=(H, W)) < 0.02] = 0.3
    err = (rng.uniform(size=(H, W)) ** 3).astype(np.float32)
    err[rng.uniform(size=(H, W)) < 0.1] = 0
    return T, err


# ---------------------------------------------------------------- tracker-side frame preprocessing
FRAMEPREP_CASES = {
    # thresholds of configs/tum (invalid_confidence_thresh 0.5) and configs/replica (0.2); depth range of base.yaml
    "tum": dict(cam="small", depth_filter=True, min_depth=0.3, max_depth=5.0, thresh=0.5),
    "nofilter_ragged": dict(cam="ragged", depth_filter=False, min_depth=0.5, max_depth=4.0, thresh=0.2),
}


def frameprep_inputs(name):
    """A noisy ray-cast depth image of the box room with holes (zero depth), as a depth sensor delivers it."""
    from rtg_slam_b200 import scene
    cfg = FRAMEPREP_CASES[name]
    cam = scene.make_camera(cfg["cam"])
    rng = np.random.default_rng(77 + cam.width)
    depth = scene.raycast_room_depth(cam).astype(np.float32)
    depth = depth + rng.normal(0, 0.004, depth.shape).astype(np.float32)
    holes = rng.uniform(size=depth.shape) < 0.03
    depth[holes] = 0
    depth[: cam.height // 6, : cam.width // 5] = 0          # a larger missing region
    K = [[cam.fx, 0.0, cam.cx], [0.0, cam.fy, cam.cy], [0.0, 0.0, 1.0]]
    return np.ascontiguousarray(depth, dtype=np.float32), K


# ----------------------------------------------------------------------------- Mapping.history_merge (mapper.py:212-250)
HISTORY_MERGE_SIZES = {"window": 777, "one": 1}


def history_merge_inputs(name):
    """Seeded (history_stat, current state) pair as local_optimize leaves them: the optimisation moved every raw parameter a
    little; a few quaternions moved a lot (the spherical branch of slerp), a few flipped sign (dot < 0), three history
    quaternions are zero (NaN dot -> linear branch). Rows whose |dot| lies within 2e-5 of the 0.9995 branch threshold are
    regenerated away from it (the branch must not depend on the last bit of a norm)."""
    P = HISTORY_MERGE_SIZES[name]
    rng = np.random.default_rng(1000 + P)
    f = np.float32
    hist_conf = rng.integers(0, 40, size=(P, 1)).astype(f)
    conf = hist_conf + rng.integers(0, 51, size=(P, 1)).astype(f)
    if P > 3:
        conf[1], hist_conf[1] = 0, 0                       # a Gaussian that never received a gradient: weight 0 / 1e-6 = 0
    hist = {"confidence": hist_conf, "xyz": rng.normal(size=(P, 3)).astype(f), "features_dc": rng.normal(size=(P, 1, 3)).astype(f),
            "features_rest": (0.1 * rng.normal(size=(P, 15, 3))).astype(f), "scaling": (rng.normal(size=(P, 3)) - 3).astype(f)}
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    hist["rotation"] = q.astype(f)
    cur = {"confidence": conf}
    for k, sig in (("xyz", 0.01), ("features_dc", 0.05), ("features_rest", 0.02), ("scaling", 0.1)):
        cur[k] = (hist[k] + sig * rng.normal(size=hist[k].shape)).astype(f)
    noise = 0.01 * rng.normal(size=(P, 4))
    big = rng.uniform(size=P) < 0.2
    noise[big] = 0.6 * rng.normal(size=(int(big.sum()), 4))
    raw = (q + noise) * rng.uniform(0.5, 2.0, size=(P, 1))    # the raw rotation is not normalised
    flip = rng.uniform(size=P) < 0.1
    raw[flip] *= -1
    cur["rotation_raw"] = raw.astype(f)
    if P > 10:
        hist["rotation"][5:8] = 0
    rn = cur["rotation_raw"].astype(np.float64)
    rn /= np.linalg.norm(rn, axis=-1, keepdims=True)
    hn = hist["rotation"].astype(np.float64)
    with np.errstate(invalid="ignore"):
        dot = np.abs((hn / np.linalg.norm(hn, axis=-1, keepdims=True) * rn).sum(-1))
    near = np.abs(dot - 0.9995) < 2e-5
    cur["rotation_raw"][near] = (hist["rotation"][near] * 1.5).astype(f)   # collinear: far inside the linear branch
    return hist, cur


def slerp_tolerance(dot, base=4e-7):
    """Per-row bound on |slerp - reference slerp| for fp32 evaluations that differ in libm / summation order: `base` on the
    linear branch, base / sin(theta_0) on the spherical one (s0, s1 = sin(.) / sin(theta_0), SLAM/utils.py:646-648)."""
    with np.errstate(invalid="ignore"):
        d = np.nan_to_num(np.abs(dot.astype(np.float64)), nan=1.0)
        lin = d > 0.9995
        s = np.sqrt(np.maximum(1.0 - np.minimum(d, 1.0) ** 2, 1e-6))
    return np.where(lin, base, base / s + base)


# ----------------------------------------------------------------------------- SSIM term (utils/loss_utils.py:40-100)
SSIM_CASES = {"ragged": (3, 37, 53), "tile": (1, 16, 16), "strip": (2, 5, 40), "smooth": (3, 48, 80)}


def ssim_inputs(name):
    """Seeded (render, frame) pair, float32 (C,H,W) in [0,1]: noise whose upper half is a perturbed copy (high SSIM), or --
    'smooth' -- low-frequency images with a black block each (sigma^2 next to C2, the ill-conditioned regime)."""
    C, H, W = SSIM_CASES[name]
    rng = np.random.default_rng(7000 + H * W)
    if name == "smooth":
        yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
        a = np.stack([0.5 + 0.4 * np.sin(6 * xx + c) * np.cos(4 * yy) for c in range(C)])
        b = np.clip(a + 0.01 * rng.normal(size=a.shape), 0, 1)
        a[:, : H // 4, : W // 3] = 0
        b[:, : H // 5, : W // 4] = 0
    else:
        a = rng.uniform(size=(C, H, W))
        b = rng.uniform(size=(C, H, W))
        b[:, : H // 2] = np.clip(a[:, : H // 2] + 0.02 * rng.normal(size=(C, H // 2, W)), 0, 1)
    return a.astype(np.float32), b.astype(np.float32)
