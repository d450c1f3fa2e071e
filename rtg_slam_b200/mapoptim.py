"""Fused parametrize + activations + attach regulariser + Adam for the Gaussian map.

What the reference does per optimisation iteration around the rasterizer (`Mapping.loss_update`,
SLAM/multiprocess/mapper.py:376-468; `GaussianPointCloud.parametrize` and the activation properties,
SLAM/gaussian_pointcloud.py:245-284,511-523,574-581):

    get_scaling = exp(_scaling); get_rotation = normalize(_rotation); get_opacity = sigmoid(_opacity)
    get_features = cat(_features_dc, _features_rest); get_normal                      (forward, ~15 launches)
    attach_loss = 1000 * (l2(_scaling[m], s0[m]) + l2(_xyz[m], x0[m]) + l2(_rotation[m], r0[m]))
    (loss + attach_loss).backward()        -> the same nodes backward, cat/slice/mask-index backward
    optimizer.step()                       -> torch.optim.Adam(l, lr=0.0, eps=1e-15): 6 groups
    _confidence[(f_dc.grad.abs() != 0).any(-1)] += 1

`MapOptimizer` keeps the raw parameters and the Adam moments, exposes the ACTIVATED tensors as autograd leaves for
`Renderer.render`, and `step()` does all of the above except the image loss in ONE kernel (`rtg_map_adam_step`):
activation backward, attach gradient, Adam, confidence, activation forward + get_normal for the next iteration.
With `rasterizer.visible_rows_only()` around the backward and `step(radii=...)`, culled rows are neither zero-filled
nor read.

    opt = MapOptimizer.from_pointcloud(pointcloud, update_args)        # replaces parametrize + torch.optim.Adam
    opt.set_attach(init_stat)                                          # mapper.py:384-401
    for it in range(iters):
        out = renderer.render(frame, opt.gaussian_data())
        loss, parts = mapping_loss(out, image_input, render_mask, ...)
        with visible_rows_only():
            loss.backward()
        opt.step(radii=out["radii"])                                   # also zero_grad(set_to_none=True)
    opt.history_merge(history_stat, self.history_merge_max_weight)     # replaces Mapping.history_merge, one kernel
    opt.write_back(pointcloud)                                         # replaces pointcloud.detach()

`history_merge(...)` (module level) is the same kernel on the reference's own tensors.

`frozen=` appends Gaussians that are rendered but not optimised -- the reference's `global_params` is
`cat(unstable (requires grad), stable (detached))`, rebuilt from both clouds' activation properties in EVERY iteration
(mapper.py:1034-1108: ~10 activation kernels over the stable map + 8 `torch.cat`). Here the stable rows are copied once
behind the optimised rows of the same buffers; `step()` only ever touches the first P rows.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import RtgHistoryMerge, RtgMapStep, check

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")  # order of parametrize's list


def _p(t):
    return None if t is None else t.data_ptr()


def _merge_call(history_stat, confidence, xyz, dc, dc_stride, rest, rest_stride, rest_width, scaling, rotation_raw, max_weight,
                device):
    P = xyz.shape[0]

    def hist(key, numel):
        t = history_stat[key]
        if not t.is_cuda or t.device != device or t.dtype != torch.float32 or t.numel() != numel:
            raise TypeError(f"history_merge: history_stat['{key}'] must be a float32 tensor with {numel} elements on {device}")
        return t.contiguous()
    keep = [hist("confidence", P), hist("xyz", 3 * P), hist("features_dc", 3 * P), hist("features_rest", rest_width * P),
            hist("scaling", 3 * P), hist("rotation", 4 * P)]
    m = RtgHistoryMerge()
    m.P, m.max_weight = P, float(max_weight)
    m.hist_confidence, m.hist_xyz, m.hist_features_dc, m.hist_features_rest, m.hist_scaling, m.hist_rotation = (_p(t) for t in keep)
    m.confidence, m.xyz, m.features_dc, m.features_rest, m.scaling, m.rotation_raw = (_p(t) for t in (confidence, xyz, dc, rest, scaling,
                                                                                                        rotation_raw))
    m.features_dc_stride, m.features_rest_stride, m.features_rest_width = int(dc_stride), int(rest_stride), int(rest_width)
    with torch.cuda.device(device):
        check(_lib.lib().rtg_map_history_merge(C.byref(m), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
              "rtg_map_history_merge")


@torch.no_grad()
def history_merge(history_stat, confidence, xyz, features_dc, features_rest, scaling, rotation, max_weight=0.5):
    """`Mapping.history_merge` (SLAM/multiprocess/mapper.py:212-250) in one kernel, IN PLACE on the raw parameter tensors of
    a GaussianPointCloud: `history_merge(history_stat, pc._confidence, pc._xyz, pc._features_dc, pc._features_rest, pc._scaling,
    pc._rotation, max_weight)`. `history_stat` is the dict local_optimize builds before the loop (mapper.py:146-155; keys
    used: confidence, xyz, features_dc, features_rest, scaling, rotation = get_rotation). Like the reference, the features
    and the scaling are merged with the first Gaussian's weight (`history_weight[0]`) and a non-positive `max_weight`
    does nothing. The tensors must be contiguous float32 CUDA tensors (they are updated through their storage)."""
    if max_weight <= 0:
        return
    dev = xyz.device
    P = xyz.shape[0]
    if P == 0:
        return
    width = features_rest.numel() // P
    for name, t, numel in (("confidence", confidence, P), ("xyz", xyz, 3 * P), ("features_dc", features_dc, 3 * P),
                           ("features_rest", features_rest, width * P), ("scaling", scaling, 3 * P), ("rotation", rotation, 4 * P)):
        if not t.is_cuda or t.device != dev or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != numel:
            raise TypeError(f"history_merge: {name} must be a contiguous float32 tensor with {numel} elements on {dev}")
    _merge_call(history_stat, confidence, xyz, features_dc, 3, features_rest, width, width, scaling, rotation, max_weight, dev)


class MapOptimizer:
    def __init__(self, xyz, features_dc, features_rest, opacity, scaling, rotation, lrs, betas=(0.9, 0.999), eps=1e-15,
                 confidence=None, attach_weight=1000.0, frozen=None):
        """Raw parameter tensors as GaussianPointCloud holds them: `_xyz (P,3)`, `_features_dc (P,1,3)`,
        `_features_rest (P,15,3)`, `_opacity (P,1)`, `_scaling (P,3)`, `_rotation (P,4)` (values are copied).
        `lrs`: dict by GROUPS (or a 6-sequence in that order). `confidence`: optional (P,) / (P,1) float tensor, updated
        in place by step(). `frozen`: optional dict of ACTIVATED tensors of S Gaussians that are rendered but not
        optimised (`Mapping.stable_params`, mapper.py:997-1022: xyz (S,3), opacity (S,1), scales (S,3), rotations (S,4),
        shs (S,16,3), normal (S,3)); they become rows P..P+S-1 of the tensors of `gaussian_data()`."""
        dev = xyz.device
        if dev.type != "cuda":
            raise TypeError("MapOptimizer: parameters must be CUDA tensors (there is no CPU path)")
        P = xyz.shape[0]
        for name, t, shape in (("xyz", xyz, (P, 3)), ("features_dc", features_dc, (P, 1, 3)), ("features_rest", features_rest, (P, 15, 3)),
                               ("opacity", opacity, (P, 1)), ("scaling", scaling, (P, 3)), ("rotation", rotation, (P, 4))):
            if t.dtype != torch.float32 or t.device != dev:
                raise TypeError(f"MapOptimizer: {name} must be float32 on {dev}")
            if tuple(t.shape) != shape:
                raise ValueError(f"MapOptimizer: {name} must have shape {shape}, got {tuple(t.shape)} "
                                 "(spherical harmonics of degree 3: 1 + 15 coefficients)")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or eps < 0:
            raise ValueError("MapOptimizer: invalid betas / eps")
        self.device, self.P = dev, P
        S = 0
        if frozen is not None:
            S = int(frozen["xyz"].shape[0])
            for name, shape in (("xyz", (S, 3)), ("opacity", (S, 1)), ("scales", (S, 3)), ("rotations", (S, 4)), ("shs", (S, 16, 3)),
                                ("normal", (S, 3))):
                t = frozen[name]
                if t.dtype != torch.float32 or t.device != dev or tuple(t.shape) != shape:
                    raise TypeError(f"MapOptimizer: frozen['{name}'] must be float32 {shape} on {dev}")
        self.S = S  # frozen rows behind the P optimised ones
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.lrs = dict(zip(GROUPS, lrs)) if not isinstance(lrs, dict) else {k: float(lrs[k]) for k in GROUPS}
        self.attach_weight = float(attach_weight)
        self.step_count = 0
        with torch.no_grad():
            # rasterizer inputs that are raw parameters themselves
            tail = lambda key: [frozen[key].detach()] if S else []
            self.xyz = torch.cat([xyz.detach()] + tail("xyz"), dim=0).contiguous().requires_grad_(True)
            self.shs = torch.cat([torch.cat([features_dc.detach(), features_rest.detach()], dim=1)] + tail("shs"),
                                 dim=0).contiguous().requires_grad_(True)  # once
            # raw parameters behind an activation
            self.opacity_raw = opacity.detach().clone().contiguous()
            self.scaling_raw = scaling.detach().clone().contiguous()
            self.rotation_raw = rotation.detach().clone().contiguous()
            # activated leaves, rewritten by every step
            self.opacity = torch.empty((P + S, 1), dtype=torch.float32, device=dev)
            self.scales = torch.empty((P + S, 3), dtype=torch.float32, device=dev)
            self.rotations = torch.empty((P + S, 4), dtype=torch.float32, device=dev)
            self.normal = torch.empty((P + S, 3), dtype=torch.float32, device=dev)
            if S:
                self.opacity[P:].copy_(frozen["opacity"])
                self.scales[P:].copy_(frozen["scales"])
                self.rotations[P:].copy_(frozen["rotations"])
                self.normal[P:].copy_(frozen["normal"])
            check(_lib.lib().rtg_map_activate(P, _p(self.scaling_raw), _p(self.rotation_raw), _p(self.opacity_raw), _p(self.scales),
                                              _p(self.rotations), _p(self.opacity), _p(self.normal), self._stream()), "rtg_map_activate")
        for t in (self.opacity, self.scales, self.rotations):
            t.requires_grad_(True)
        self.state = {k: (torch.zeros_like(t), torch.zeros_like(t)) for k, t in
                      (("xyz", self.xyz.detach()[:P]), ("sh", self.shs.detach()[:P]), ("opacity", self.opacity_raw),
                       ("scaling", self.scaling_raw), ("rotation", self.rotation_raw))}
        if confidence is not None and (confidence.dtype != torch.float32 or confidence.numel() != P or not confidence.is_contiguous()
                                       or confidence.device != dev):
            raise TypeError("MapOptimizer: confidence must be a contiguous float32 tensor with one element per Gaussian")
        self.confidence = confidence
        self._attach = None
        self._st = None

    # ------------------------------------------------------------------ construction from / write-back to the reference's store
    @classmethod
    def from_pointcloud(cls, pc, update_args, lr_scale=None, **kw):
        """`pc`: a GaussianPointCloud; `update_args`: the namespace parametrize reads (position_lr, feature_lr, opacity_lr,
        scaling_lr, rotation_lr). `lr_scale`: optional dict of per-group multipliers (global optimisation scales xyz /
        scaling / rotation, mapper.py:606-615)."""
        lrs = {"xyz": update_args.position_lr, "f_dc": update_args.feature_lr, "f_rest": update_args.feature_lr / 20.0,
               "opacity": update_args.opacity_lr, "scaling": update_args.scaling_lr, "rotation": update_args.rotation_lr}
        for k, v in (lr_scale or {}).items():
            lrs[k] *= v
        conf = getattr(pc, "_confidence", None)
        return cls(pc._xyz, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling, pc._rotation, lrs,
                   confidence=None if conf is None else conf.view(-1), **kw)

    def write_back(self, pc):
        """The optimised raw parameters into a GaussianPointCloud, detached (what `pointcloud.detach()` leaves behind,
        gaussian_pointcloud.py:306-313). `_features_dc` / `_features_rest` become views of one (P,16,3) block."""
        pc._xyz = self.xyz.detach()[:self.P]
        pc._features_dc = self.features_dc
        pc._features_rest = self.features_rest
        pc._opacity = self.opacity_raw
        pc._scaling = self.scaling_raw
        pc._rotation = self.rotation_raw
        return pc

    # ------------------------------------------------------------------ history merge (mapper.py:146-155,210-250)
    def history_snapshot(self):
        """The `history_stat` dict `Mapping.local_optimize` builds before its loop (mapper.py:146-155) from the optimiser's
        current state: clones of the raw parameters, the activated rotation and the confidence."""
        if self.confidence is None:
            raise RuntimeError("MapOptimizer.history_snapshot: no confidence tensor was given to the optimiser")
        return {"opacity": self.opacity_raw.clone(), "confidence": self.confidence.detach().clone().view(-1, 1),
                "xyz": self.xyz.detach()[:self.P].clone(), "features_dc": self.features_dc.clone(),
                "features_rest": self.features_rest.clone(), "scaling": self.scaling_raw.clone(),
                "rotation": self.rotations.detach()[:self.P].clone(), "rotation_raw": self.rotation_raw.clone()}

    @torch.no_grad()
    def history_merge(self, history_stat, max_weight=0.5):
        """`Mapping.history_merge` (mapper.py:212-250) on the optimiser's own tensors, in place, followed by the activation
        forward so that `gaussian_data()` is consistent again. Call it after the loop, before `write_back`."""
        if max_weight <= 0 or self.P == 0:
            return
        if self.confidence is None:
            raise RuntimeError("MapOptimizer.history_merge: no confidence tensor was given to the optimiser")
        sh = self.shs.detach()[:self.P]
        _merge_call(history_stat, self.confidence, self.xyz.detach()[:self.P], sh, 48, sh.view(-1)[3:], 48, 45, self.scaling_raw,
                    self.rotation_raw, max_weight, self.device)
        check(_lib.lib().rtg_map_activate(self.P, _p(self.scaling_raw), _p(self.rotation_raw), _p(self.opacity_raw), _p(self.scales),
                                          _p(self.rotations), _p(self.opacity), _p(self.normal), self._stream()), "rtg_map_activate")

    # ------------------------------------------------------------------ views
    @property
    def features_dc(self):
        return self.shs.detach()[:self.P, :1]

    @property
    def features_rest(self):
        return self.shs.detach()[:self.P, 1:]

    def gaussian_data(self):
        """The dict `Renderer.render` takes (SLAM/render.py:93-98; what Mapping.get_render_output builds from the
        pointcloud's get_* properties; with `frozen`, the reference's `global_params`: optimised rows first, frozen rows
        behind them). The tensors are autograd leaves: backward leaves their `.grad` for step()."""
        return {"xyz": self.xyz, "opacity": self.opacity, "scales": self.scales, "rotations": self.rotations, "shs": self.shs,
                "normal": self.normal}

    # ------------------------------------------------------------------ attach regulariser
    def set_attach(self, init_stat):
        """`init_stat`: dict with "opacity" (raw), "scaling", "xyz", "rotation_raw" as mapper.py:617-622 builds it; the mask
        is sigmoid(opacity) < 0.9 (mapper.py:384-385). One host read of the mask count, once per optimisation call. None
        removes the term."""
        self._st = None
        if init_stat is None:
            self._attach = None
            return
        mask = (torch.sigmoid(init_stat["opacity"]) < 0.9).reshape(-1)
        if mask.numel() != self.P:
            raise ValueError("MapOptimizer.set_attach: init_stat has a different number of Gaussians")
        count = int(mask.sum().item())
        self._attach = None if count == 0 else dict(
            mask=mask.to(torch.uint8).contiguous(), count=count, xyz0=init_stat["xyz"].detach().float().contiguous(),
            scaling0=init_stat["scaling"].detach().float().contiguous(), rotation0=init_stat["rotation_raw"].detach().float().contiguous())

    def attach_loss(self):
        """Value of the regulariser (for reporting; its gradient is applied inside step())."""
        if self._attach is None:
            return torch.zeros((), device=self.device)
        m = self._attach["mask"].bool()

        def l2(a, b):
            return ((a[m] - b[m]) ** 2).mean()
        with torch.no_grad():
            return self.attach_weight * (l2(self.scaling_raw, self._attach["scaling0"]) + l2(self.xyz.detach()[:self.P], self._attach["xyz0"])
                                         + l2(self.rotation_raw, self._attach["rotation0"]))

    # ------------------------------------------------------------------ the step
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _step_struct(self):
        """The argument block of rtg_map_adam_step; everything but the gradients, radii and the step number is fixed
        between set_attach calls, so it is filled once."""
        if self._st is None:
            st = RtgMapStep()
            st.P = self.P
            st.xyz, st.sh, st.opacity_raw = _p(self.xyz), _p(self.shs), _p(self.opacity_raw)
            st.scaling_raw, st.rotation_raw = _p(self.scaling_raw), _p(self.rotation_raw)
            for key in ("xyz", "sh", "opacity", "scaling", "rotation"):
                m, v = self.state[key]
                setattr(st, "m_" + key, _p(m))
                setattr(st, "v_" + key, _p(v))
            if self._attach is not None:
                a = self._attach
                st.attach_mask, st.xyz0, st.scaling0, st.rotation0 = _p(a["mask"]), _p(a["xyz0"]), _p(a["scaling0"]), _p(a["rotation0"])
                st.attach_weight, st.attach_count = self.attach_weight, a["count"]
            st.lr_xyz, st.lr_f_dc, st.lr_f_rest = self.lrs["xyz"], self.lrs["f_dc"], self.lrs["f_rest"]
            st.lr_opacity, st.lr_scaling, st.lr_rotation = self.lrs["opacity"], self.lrs["scaling"], self.lrs["rotation"]
            st.beta1, st.beta2, st.eps = self.betas[0], self.betas[1], self.eps
            st.scales_out, st.rotations_out, st.opacities_out = _p(self.scales), _p(self.rotations), _p(self.opacity)
            st.normal_out, st.confidence = _p(self.normal), _p(self.confidence)
            self._st = st
        return self._st

    def zero_grad(self, set_to_none=True):
        for t in (self.xyz, self.shs, self.opacity, self.scales, self.rotations):
            if t.grad is not None:
                if set_to_none:
                    t.grad = None
                else:
                    t.grad.zero_()

    @torch.no_grad()
    def step(self, radii=None, zero_grad=True):
        """One Adam step from the `.grad` of the five rasterizer inputs. `radii`: the forward's radii (int32, P) -- required
        when the backward ran under `rasterizer.visible_rows_only()`: rows with radii <= 0 then count as zero gradient and
        are not read. Rewrites the activated tensors (and `normal`) in place and, by default, drops the gradients."""
        grads = {}
        for name, t in (("xyz", self.xyz), ("shs", self.shs), ("opacity", self.opacity), ("scales", self.scales),
                        ("rotations", self.rotations)):
            g = t.grad
            if g is None:
                raise RuntimeError(f"MapOptimizer.step: '{name}' has no gradient (run loss.backward() on a render of gaussian_data())")
            if g.dtype != torch.float32 or g.device != self.device:
                raise TypeError(f"MapOptimizer.step: gradient of '{name}' must be float32 on {self.device}")
            grads[name] = g if g.is_contiguous() else g.contiguous()
        if radii is not None and (radii.dtype != torch.int32 or radii.numel() != self.P + self.S or radii.device != self.device
                                  or not radii.is_contiguous()):
            raise TypeError("MapOptimizer.step: radii must be the rasterizer's contiguous int32 output (one entry per rendered Gaussian)")
        self.step_count += 1
        st = self._step_struct()
        st.step = self.step_count
        st.g_means3D, st.g_sh, st.g_opacity = _p(grads["xyz"]), _p(grads["shs"]), _p(grads["opacity"])
        st.g_scales, st.g_rotations = _p(grads["scales"]), _p(grads["rotations"])
        st.radii = _p(radii)
        with torch.cuda.device(self.device):
            check(_lib.lib().rtg_map_adam_step(C.byref(st), self._stream()), "rtg_map_adam_step")
        if zero_grad:
            self.zero_grad(set_to_none=True)
