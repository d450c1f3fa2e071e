"""Multi-GPU plumbing of the hot path: one process per GPU over torch.distributed (NCCL on the B200s, gloo in the
CPU tests).

What shards and what does not (DESIGN.md "Multi-GPU"):
* the Gaussian map is REPLICATED: rank 0 owns the authoritative copy and broadcasts it (`broadcast_map`) when it
  changes (once per SLAM frame, when Gaussians are added / removed);
* FRAMES shard: the optimisation window of `Mapping.local_optimize` (SLAM/multiprocess/mapper.py:143-210) holds
  several keyframes, and rendering + back-propagating one frame is independent of the others -- `shard_frames`
  gives every rank its frames, no collective is on that path;
* the per-Gaussian gradients of the frames processed in one step are summed with ONE all-reduce over a single flat
  buffer (`FlatGrads`) before the (identical, replicated) Adam step -- the only exchange step of the loop;
* TILES shard (single-frame strong scaling, SURVEY.md section 8(e) "literal first version"): compositing is
  independent per 16x16 tile, so `TileShard` gives every rank an interleaved (or load-balanced) subset of the tiles
  as the rasterizer's own `tile_mask` (RAST/.../__init__.py:210, forward.cu:660); every rank renders and
  back-propagates only its tiles against the replicated map, the image pieces are disjoint (`TileShard.gather`
  reassembles them when the full image is needed; the loss can stay shard-wise) and the per-Gaussian gradients are
  partial sums, completed either by the same flat all-reduce (236 B / Gaussian) or -- `TileShard.exchange_records` --
  by summing the 64-byte gradient records between the two halves of the backward (rtg_splat_backward_render /
  _finish), after which every rank computes the identical full gradient itself;
* ICP is a 27-number reduction per iteration at <= 0.8 Mpx: replicas only, never sharded.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

# (name, trailing shape) of the gradient tensors the rasterizer returns, in flat-buffer order
GRAD_LAYOUT = (("means3D", (3,)), ("shs", (16, 3)), ("opacities", (1,)), ("scales", (3,)), ("rotations", (4,)))


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def broadcast_map(tensors: Dict[str, torch.Tensor], src: int = 0) -> Dict[str, torch.Tensor]:
    """In-place broadcast of the Gaussian map (all ranks must pass tensors of the same shapes)."""
    if world() > 1:
        for k in sorted(tensors):
            dist.broadcast(tensors[k], src=src)
    return tensors


def shard_frames(n_frames: int, world_size: int | None = None, r: int | None = None) -> List[int]:
    """Frames of the window owned by rank r: round robin, so that ranks differ by at most one frame."""
    w = world() if world_size is None else world_size
    r = rank() if r is None else r
    return list(range(r, n_frames, w))


class FlatGrads:
    """One contiguous fp32 buffer holding all per-Gaussian gradient tensors, with typed views into it. Inside
    `rasterizer.grad_buffers(flat.views)` the rasterizer backward writes straight into the views; `allreduce` then needs
    a single collective and no packing (`accumulate` is the packing path for gradients produced elsewhere)."""

    def __init__(self, P: int, device, sh_coeffs: int = 16):
        self.P = P
        layout = [(n, (sh_coeffs, 3) if n == "shs" else s) for n, s in GRAD_LAYOUT]
        sizes = [P * int(torch.tensor(s).prod()) for _, s in layout]
        # keep every view 16-byte aligned (the kernels use 128-bit stores)
        self.offsets, off = [], 0
        for sz in sizes:
            self.offsets.append(off)
            off += (sz + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = {n: self.flat[o:o + sz].view(P, *s) for (n, s), o, sz in zip(layout, self.offsets, sizes)}

    def zero_(self):
        self.flat.zero_()
        return self

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        for n, v in self.views.items():
            v.add_(grads[n].view_as(v))
        return self

    def allreduce(self, average: bool = False, async_op: bool = False):
        if average and async_op:
            raise ValueError("FlatGrads.allreduce: average=True needs the result, i.e. async_op=False (divide after wait())")
        if world() == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if average:
            self.flat.div_(world())
        return work


class TileShard:
    """Ownership of the 16x16 tiles of one frame by the ranks of the job.

    `mask` (th, tw) int32 is what the rank passes to the rasterizer as `tile_mask` (AND-ed with the caller's own mask,
    e.g. the mapper's transmission / colour-error masks); `pixel_mask` (H, W) bool selects the pixels the rank owns.
    Ownership is a pure function of (tile grid, world size, weights): every rank computes the same table without
    communicating. Without weights tiles are dealt round robin along the row-major tile index (neighbouring tiles
    have similar list lengths, so this already balances well); with per-tile `weights` (e.g. the list lengths of
    the previous iteration) tiles are dealt longest-first to the least loaded rank (LPT)."""

    TILE = 16

    def __init__(self, H: int, W: int, world_size: int | None = None, r: int | None = None, base_mask: torch.Tensor | None = None,
                 weights: torch.Tensor | None = None, device="cpu", bands: bool = False):
        self.H, self.W = H, W
        self.world = world() if world_size is None else world_size
        self.rank = rank() if r is None else r
        th, tw = (H + self.TILE - 1) // self.TILE, (W + self.TILE - 1) // self.TILE
        self.tile_grid = (th, tw)
        if bands:
            # contiguous bands of tile rows with (nearly) equal total weight: row r goes to the rank whose share of the
            # cumulative weight its midpoint falls into
            row_w = torch.ones(th, dtype=torch.float64) if weights is None else \
                weights.detach().to("cpu", torch.float64).reshape(th, tw).sum(1).clamp_min(1e-12)
            mid = torch.cumsum(row_w, 0) - 0.5 * row_w
            own_row = torch.clamp((mid / row_w.sum() * self.world).floor().long(), 0, self.world - 1)
            self.owner = own_row[:, None].expand(th, tw).contiguous()
            mine_rows = torch.nonzero(own_row == self.rank).flatten()
            self.rows = (int(mine_rows[0]), int(mine_rows[-1]) + 1) if mine_rows.numel() else (0, 0)
        else:
            self.owner = self._owners(th * tw, self.world, weights).view(th, tw)
            self.rows = (0, th)
        mine = self.owner == self.rank
        if base_mask is not None:
            mine = mine & (base_mask.to("cpu") != 0)
        self.mask = mine.to(torch.int32).contiguous().to(device)
        self.pixel_mask = mine.repeat_interleave(self.TILE, 0).repeat_interleave(self.TILE, 1)[:H, :W].contiguous().to(device)

    @staticmethod
    def _owners(n_tiles: int, w: int, weights: torch.Tensor | None) -> torch.Tensor:
        if weights is None:
            return torch.arange(n_tiles, dtype=torch.int64) % w
        wt = weights.detach().to("cpu", torch.float64).reshape(-1)
        assert wt.numel() == n_tiles
        order = torch.argsort(wt, descending=True, stable=True).tolist()
        load = [0.0] * w
        owner = torch.empty(n_tiles, dtype=torch.int64)
        for t in order:
            r = min(range(w), key=lambda k: (load[k], k))
            owner[t] = r
            load[r] += float(wt[t])
        return owner

    def exchange_records(self):
        """Context manager: while active, every rasterizer backward sums its (P, 16) gradient records over the ranks
        (one all-reduce of 64 B per Gaussian, between the compositing backward and the per-Gaussian backward) instead of
        leaving partial dense gradients to be all-reduced afterwards (236 B per Gaussian). All ranks then compute the
        identical, complete gradient; no further collective is needed before the replicated optimizer step."""
        import contextlib

        from . import rasterizer

        @contextlib.contextmanager
        def cm():
            def hook(rec):
                if self.world > 1:
                    dist.all_reduce(rec, op=dist.ReduceOp.SUM)
            prev = rasterizer.set_grad_record_hook(hook)
            try:
                yield self
            finally:
                rasterizer.set_grad_record_hook(prev)
        return cm()

    def gather(self, img: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
        """Full image from the ranks' disjoint pieces: pixels a rank does not own are replaced by zero, one SUM
        all-reduce, then `fill` is restored where nobody rendered (the rasterizer writes 0 / T=1 in tiles that are
        masked out, so pass fill=1 for the transmittance map)."""
        own = self.pixel_mask
        piece = torch.where(own, img, torch.zeros((), dtype=img.dtype, device=img.device))
        if self.world > 1:
            dist.all_reduce(piece, op=dist.ReduceOp.SUM)
            covered = own.to(torch.int32)
            dist.all_reduce(covered, op=dist.ReduceOp.SUM)
        else:
            covered = own.to(torch.int32)
        if fill != 0.0:
            piece = torch.where(covered > 0, piece, torch.full((), fill, dtype=img.dtype, device=img.device))
        return piece


def assert_replicas_equal(tensors: Sequence[torch.Tensor], atol: float = 0.0) -> None:
    """Debug check used by the tests: every rank holds the same values."""
    if world() == 1:
        return
    for t in tensors:
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        if not torch.allclose(ref, t, atol=atol, rtol=0):
            raise AssertionError(f"rank {rank()} diverged from rank 0 (max abs diff {(ref - t).abs().max().item():.3e})")


# ----------------------------------------------------------------------------- Gaussian-sharded rendering
def shard_range(P: int, world_size: int, r: int):
    """[p_begin, p_end) of the map owned by rank r: equal slices (the exchanges are all-gather / reduce-scatter of equal
    pieces, so P must be a multiple of the world size: pad the map with culled Gaussians, e.g. behind the camera)."""
    if P % world_size != 0:
        raise ValueError(f"P = {P} must be a multiple of the world size {world_size} (pad the map)")
    n = P // world_size
    return r * n, (r + 1) * n


class _Collectives:
    """all-gather of equal row slices / reduce-scatter of equal row slices; over torch.distributed when a group is up,
    emulated in-process for the single-GPU tests (`peers`: the other emulated ranks share the buffers)."""

    def __init__(self, world_size, r):
        self.world, self.rank = world_size, r

    def all_gather_rows(self, full: torch.Tensor):
        """`full` (P, ...) with this rank's slice filled in: on return every slice holds its owner's data."""
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            return
        n = full.shape[0] // self.world
        mine = full[self.rank * n:(self.rank + 1) * n]
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(full, mine)  # in place: `mine` is this rank's slice of `full`
        else:  # gloo (CPU tests): list form
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.contiguous())
            for k, p in enumerate(parts):
                full[k * n:(k + 1) * n].copy_(p)

    def reduce_scatter_rows(self, full: torch.Tensor, out: torch.Tensor):
        """out (P / world, ...) = sum over ranks of their `full`[own slice of this rank]."""
        n = full.shape[0] // self.world
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            out.copy_(full[self.rank * n:(self.rank + 1) * n])
            return
        if dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM)
        else:
            tmp = full.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
            out.copy_(tmp[self.rank * n:(self.rank + 1) * n])


class GaussianShard:
    """One rank's part of a Gaussian-sharded frame (BASELINE configs[4]).

    The rank owns the parameters (and gradients, optimizer state) of the Gaussians [p_begin, p_end) and a subset of the
    tiles (`TileShard`). Forward: preprocess of the owned Gaussians -> all-gather of the per-Gaussian records (80 B +
    radius per Gaussian) -> binning + compositing of the owned tiles against all records. Backward: compositing backward
    of the owned tiles -> reduce-scatter of the 64-byte gradient records to the owners -> per-Gaussian backward of the
    owned Gaussians. No parameter is ever replicated; the only full-size buffers are the records.

        sh = GaussianShard(P, H, W, device)
        out = sh.forward(raster_settings, means3D, opacities, shs, scales, rotations)      # owned rows
        grads = sh.backward(dL_dcolor, dL_ddepth)                                           # owned rows
    """

    def __init__(self, P: int, H: int, W: int, device, world_size: int | None = None, r: int | None = None, tile_weights=None,
                 collectives=None):
        from . import _lib
        self.P, self.H, self.W = P, H, W
        self.world = world() if world_size is None else world_size
        self.rank = rank() if r is None else r
        self.device = torch.device(device)
        self.p_begin, self.p_end = shard_range(P, self.world, self.rank)
        # tiles: a band of tile rows per rank (contiguous, so that the binning passes can clip every Gaussian's rectangle
        # to the band before expanding it); `tile_weights` (per tile, e.g. the previous frame's list lengths) balances the bands
        self.tiles = TileShard(H, W, self.world, self.rank, weights=tile_weights, device=self.device, bands=True)
        self.row_begin, self.row_end = self.tiles.rows
        self.coll = collectives if collectives is not None else _Collectives(self.world, self.rank)
        self.r_cap = 1 << 16
        self._lib = _lib
        L = _lib.lib()
        import ctypes as C
        offs = [C.c_size_t() for _ in range(4)]
        _lib.check(L.rtg_splat_geom_layout(P, *[C.byref(o) for o in offs]), "rtg_splat_geom_layout")
        self._geom_off = [o.value for o in offs]
        gb, ib, bb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.check(L.rtg_splat_workspace_bytes(P, H, W, self.r_cap, C.byref(gb), C.byref(ib), C.byref(bb)), "rtg_splat_workspace_bytes")
        self.geom = torch.empty(gb.value, dtype=torch.uint8, device=self.device)
        self.img = torch.empty(ib.value, dtype=torch.uint8, device=self.device)
        self.bin = None
        self.radii = torch.zeros(P, dtype=torch.int32, device=self.device)
        self.rec_full = torch.zeros((P, 16), dtype=torch.float32, device=self.device)   # gradient records, zero between uses
        self.rec_own = torch.empty((self.p_end - self.p_begin, 16), dtype=torch.float32, device=self.device)
        self.counters = torch.zeros(_lib.RTG_CNT_WORDS, dtype=torch.int32, device=self.device)
        self.pinned = torch.zeros(_lib.RTG_CNT_WORDS, dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event() if self.device.type == "cuda" else None
        if self.event is not None:
            with torch.cuda.device(self.device):
                self.event.record()
        self._saved = None

    # typed views of the record arrays inside the geometry workspace (what the all-gather moves)
    def _record_views(self):
        P = self.P
        o_splat, o_rgb, o_hit, _ = self._geom_off
        g = self.geom
        return [g[o_splat:o_splat + 32 * P].view(torch.float32).view(P, 8), g[o_rgb:o_rgb + 16 * P].view(torch.float32).view(P, 4),
                g[o_hit:o_hit + 32 * P].view(torch.float32).view(P, 8)]

    def exchange_bytes(self):
        """bytes received per rank and step: records all-gather + gradient-record reduce-scatter"""
        frac = (self.world - 1) / self.world
        return {"all_gather": int((32 + 16 + 32 + 4) * self.P * frac), "reduce_scatter": int(64 * self.P * frac)}

    def forward(self, rs, means3D, opacities, shs, scales, rotations, tile_mask=None, staged=False):
        import ctypes as C
        from .rasterizer import _make_view, _f32_cuda
        L, lib = self._lib.lib(), self._lib
        dev = self.device
        n = self.p_end - self.p_begin
        for t, cols in ((means3D, 3), (opacities, 1), (scales, 3), (rotations, 4)):
            if t.shape[0] != n:
                raise ValueError(f"parameter tensors must hold the {n} owned rows")
        means3D, opacities, shs, scales, rotations = (_f32_cuda(t, nm, dev) for t, nm in (
            (means3D, "means3D"), (opacities, "opacities"), (shs, "shs"), (scales, "scales"), (rotations, "rotations")))
        M = shs.shape[1]
        view, keep = _make_view(rs, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        H, W, P = self.H, self.W, self.P
        mask = self.tiles.mask if tile_mask is None else (self.tiles.mask * (tile_mask.to(dev) != 0).to(torch.int32)).contiguous()
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        outs = [torch.empty((3, H, W), **f32), torch.empty((1, H, W), **f32), torch.empty((1, H, W), **i32), torch.empty((1, H, W), **i32),
                torch.empty((1, H, W), **f32), torch.empty((1, H, W), **f32), torch.empty((1, H, W), **f32)]
        p = lambda t: C.c_void_p(t.data_ptr())
        gb, ib, bb = C.c_size_t(), C.c_size_t(), C.c_size_t()

        def preprocess():
            lib.check(L.rtg_splat_workspace_bytes(P, H, W, self.r_cap, C.byref(gb), C.byref(ib), C.byref(bb)), "rtg_splat_workspace_bytes")
            if self.bin is None or self.bin.numel() < bb.value:
                self.bin = torch.empty(bb.value, dtype=torch.uint8, device=dev)
            lib.check(L.rtg_splat_forward_preprocess(C.byref(view), P, self.p_begin, self.p_end, M, p(means3D), p(shs), None, p(opacities),
                                                     p(scales), p(rotations), None, p(self.geom), p(self.bin), self.r_cap,
                                                     p(self.radii), C.c_void_p(stream)), "rtg_splat_forward_preprocess")

        def render():
            """False if the binning buffer was too small (the caller grows it, preprocesses and exchanges again: the
            visible-list counter lives in the binning workspace, which moves)."""
            lib.check(L.rtg_splat_forward_render(C.byref(view), P, p(mask), p(self.geom), p(self.img), p(self.bin), self.r_cap,
                                                 *[p(o) for o in outs], p(self.radii), p(self.counters), C.c_void_p(self.pinned.data_ptr()),
                                                 C.c_void_p(self.event.cuda_event), self.row_begin, self.row_end, C.c_void_p(stream)),
                      "rtg_splat_forward_render")
            self.event.synchronize()  # the scan kernel's counters (the compositing is still running)
            self.num_rendered, overflow = int(self.pinned[0]), int(self.pinned[2])
            need = max(self.num_rendered, int(self.pinned[4]))  # with the tile buckets padded to 16-byte boundaries
            if overflow:
                self.r_cap = int(need * 1.5) + 4096
                return False
            self.r_cap = max(self.r_cap, int(need * 1.25) + 4096)
            return True

        self._saved = (view, keep, M, means3D, shs, scales, rotations, outs, mask)
        names = ("color", "depth", "hit_color", "hit_depth", "hit_color_weight", "hit_depth_weight", "T_map")
        res = dict(zip(names, outs))
        res["radii"] = self.radii
        if staged:  # the caller performs the exchange itself (single-process emulation of several ranks in the tests)
            return res, preprocess, render
        while True:
            preprocess()
            self.exchange_records_forward()
            if render():
                return res

    def exchange_records_forward(self):
        for v in self._record_views():
            self.coll.all_gather_rows(v)
        self.coll.all_gather_rows(self.radii)

    def backward(self, dL_dcolor, dL_ddepth, staged=False):
        import ctypes as C
        from .rasterizer import _f32_cuda
        L, lib = self._lib.lib(), self._lib
        dev = self.device
        view, keep, M, means3D, shs, scales, rotations, outs, mask = self._saved
        n = self.p_end - self.p_begin
        f32 = dict(dtype=torch.float32, device=dev)
        g = dict(means3D=torch.empty((n, 3), **f32), shs=torch.empty((n, M, 3), **f32), opacities=torch.empty((n, 1), **f32),
                 scales=torch.empty((n, 3), **f32), rotations=torch.empty((n, 4), **f32))
        gc, gd = _f32_cuda(dL_dcolor, "dL_dcolor", dev), _f32_cuda(dL_ddepth, "dL_ddepth", dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())

        def args(rec_ptr):
            return (self.p_begin, self.p_end, C.byref(view), self.P, M, p(means3D), p(shs), None, p(scales), p(rotations), None,
                    p(self.radii), p(self.geom), p(self.img), p(self.bin), self.r_cap, p(self.counters), p(outs[6]), p(outs[3]), p(gc), p(gd),
                    rec_ptr, p(g["means3D"]), p(g["shs"]), None, p(g["opacities"]), p(g["scales"]), p(g["rotations"]), None, None, stream)
        def render():
            lib.check(L.rtg_splat_backward_render_shard(*args(p(self.rec_full))), "rtg_splat_backward_render_shard")

        def finish():
            self.rec_full.zero_()
            lib.check(L.rtg_splat_backward_finish_shard(*args(C.c_void_p(self.rec_own.data_ptr() - self.p_begin * 64))),
                      "rtg_splat_backward_finish_shard")
            return g
        if staged:
            return render, finish
        render()
        self.coll.reduce_scatter_rows(self.rec_full, self.rec_own)
        return finish()
