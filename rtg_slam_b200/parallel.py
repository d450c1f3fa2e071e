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
    """One contiguous fp32 buffer holding all per-Gaussian gradient tensors, with typed views into it. The
    rasterizer backward writes into the views; `allreduce` then needs a single collective and no packing."""

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
        if world() == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if average and not async_op:
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
                 weights: torch.Tensor | None = None, device="cpu"):
        self.H, self.W = H, W
        self.world = world() if world_size is None else world_size
        self.rank = rank() if r is None else r
        th, tw = (H + self.TILE - 1) // self.TILE, (W + self.TILE - 1) // self.TILE
        self.tile_grid = (th, tw)
        self.owner = self._owners(th * tw, self.world, weights).view(th, tw)
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
