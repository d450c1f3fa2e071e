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


def assert_replicas_equal(tensors: Sequence[torch.Tensor], atol: float = 0.0) -> None:
    """Debug check used by the tests: every rank holds the same values."""
    if world() == 1:
        return
    for t in tensors:
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        if not torch.allclose(ref, t, atol=atol, rtol=0):
            raise AssertionError(f"rank {rank()} diverged from rank 0 (max abs diff {(ref - t).abs().max().item():.3e})")
