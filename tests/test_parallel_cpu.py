"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in rtg_slam_b200/parallel.py: map broadcast, frame
sharding, the flat gradient all-reduce, and that a replicated optimizer step keeps the ranks identical. Per-rank
gradients come from the CPU oracle (different camera per frame), so the test also checks that the reduced gradient
equals the single-process sum over the window."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtg_slam_b200 import parallel, scene

WINDOW = 3  # keyframes in the optimisation window


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame_grads(g, frame):
    from oracle.splat_oracle import OracleRender
    cam = scene.make_camera("tiny", c2w=scene.small_pose((0.5 * frame, -0.4 * frame, 0.2 * frame), (0.01 * frame, 0.0, 0.01 * frame)))
    o = OracleRender(cam, g, precision="f32")
    gr = o.backward(*scene.upstream_grads(cam, seed=5 + frame))
    o.close()
    return {k: torch.from_numpy(gr[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations")}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 400
        g_np = scene.surfel_room(P, seed=3) if rank == 0 else None
        shapes = {"xyz": (P, 3), "opacity": (P, 1), "scales": (P, 3), "rotations": (P, 4), "shs": (P, 16, 3), "normal": (P, 3)}
        t = {k: (torch.from_numpy(g_np[k]) if rank == 0 else torch.full(shapes[k], float("nan"))) for k in shapes}
        parallel.broadcast_map(t)
        parallel.assert_replicas_equal(list(t.values()))
        g = {k: v.numpy() for k, v in t.items()}
        mine = parallel.shard_frames(WINDOW)
        fg = parallel.FlatGrads(P, "cpu")
        for f in mine:
            fg.accumulate(_frame_grads(g, f))
        fg.allreduce()
        # replicated optimizer step on the reduced gradient
        params = [t["xyz"].clone().requires_grad_(True), t["shs"].clone().requires_grad_(True)]
        opt = torch.optim.Adam([{"params": [params[0]], "lr": 1e-3}, {"params": [params[1]], "lr": 5e-4}], lr=0.0, eps=1e-15)
        params[0].grad = fg.views["means3D"].clone()
        params[1].grad = fg.views["shs"].clone()
        opt.step()
        parallel.assert_replicas_equal([p.detach() for p in params])
        q.put((rank, mine, {k: v.clone().numpy() for k, v in fg.views.items()}))
    finally:
        dist.destroy_process_group()


def test_world2_broadcast_shard_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    assert sorted(res[0][1] + res[1][1]) == list(range(WINDOW))  # every frame exactly once
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1
    g = scene.surfel_room(400, seed=3)
    total = None
    for f in range(WINDOW):
        gr = _frame_grads(g, f)
        total = gr if total is None else {k: total[k] + gr[k] for k in gr}
    for r in res:
        for k in total:
            ref = total[k].numpy().reshape(r[2][k].shape)
            assert np.allclose(r[2][k], ref, rtol=1e-5, atol=1e-9), k
    assert any(np.abs(v).max() > 0 for v in res[0][2].values())


def test_shard_frames_and_flat_layout_single_process():
    assert parallel.shard_frames(7, 3, 0) == [0, 3, 6] and parallel.shard_frames(7, 3, 2) == [2, 5]
    assert parallel.shard_frames(2, 8, 5) == []
    fg = parallel.FlatGrads(5, "cpu")
    assert fg.views["shs"].shape == (5, 16, 3) and fg.views["rotations"].shape == (5, 4)
    for v in fg.views.values():
        assert v.data_ptr() % 16 == 0
    fg.views["opacities"].fill_(2.0)
    assert fg.flat.sum().item() == 10.0
    assert fg.allreduce() is None  # no process group: no-op


# ---------------------------------------------------------------------------------------------- tile sharding
def _tile_worker(rank, world, port, q):
    """One frame, tiles dealt to the ranks: rendering (oracle) only the own tiles, the gathered image equals the
    single-process image and the all-reduced partial gradients equal the single-process gradients."""
    from oracle.splat_oracle import OracleRender
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 600
        g = scene.surfel_room(P, seed=11)
        cam = scene.make_camera("tiny")
        th, tw = cam.tile_grid
        base = torch.ones(th, tw, dtype=torch.int32)
        base[0, 0] = 0  # a tile the caller itself masks out stays masked on every rank
        shard = parallel.TileShard(cam.height, cam.width, base_mask=base)
        o = OracleRender(cam, g, tile_mask=shard.mask.numpy(), precision="f32")
        color, depth, T = (torch.from_numpy(np.array(a)) for a in (o.color, o.depth, o.T_map))
        gc, gd = scene.upstream_grads(cam, seed=5)
        gr = o.backward(gc, gd)
        o.close()
        fg = parallel.FlatGrads(P, "cpu")
        fg.accumulate({k: torch.from_numpy(gr[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations")})
        fg.allreduce()
        full = (shard.gather(color), shard.gather(depth), shard.gather(T, fill=1.0))
        q.put((rank, int(shard.mask.sum()), [f.numpy() for f in full], {k: v.clone().numpy() for k, v in fg.views.items()}))
    finally:
        dist.destroy_process_group()


def test_world2_tile_sharded_frame():
    from oracle.splat_oracle import OracleRender
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = scene.surfel_room(600, seed=11)
    cam = scene.make_camera("tiny")
    th, tw = cam.tile_grid
    base = np.ones((th, tw), np.int32)
    base[0, 0] = 0
    assert res[0][1] + res[1][1] == th * tw - 1 and abs(res[0][1] - res[1][1]) <= 1
    o = OracleRender(cam, g, tile_mask=base, precision="f32")
    gr = o.backward(*scene.upstream_grads(cam, seed=5))
    for r in res:  # every rank ends up with the full image and the full gradient
        assert np.array_equal(r[2][0], o.color) and np.array_equal(r[2][1], o.depth) and np.array_equal(r[2][2], o.T_map)
        for k in ("means3D", "shs", "opacities", "scales", "rotations"):
            ref = gr[k].reshape(r[3][k].shape)
            assert np.allclose(r[3][k], ref, rtol=1e-4, atol=1e-9 + 1e-6 * np.abs(ref).max()), k
    assert np.abs(gr["means3D"]).max() > 0
    o.close()


def test_tile_shard_tables():
    H, W = 680, 1200
    shards = [parallel.TileShard(H, W, 4, r) for r in range(4)]
    total = sum(s.mask for s in shards)
    assert torch.equal(total, torch.ones_like(total))  # a partition
    assert max(int(s.mask.sum()) for s in shards) - min(int(s.mask.sum()) for s in shards) <= 1
    union = torch.zeros(H, W, dtype=torch.int32)
    for s in shards:
        union += s.pixel_mask.to(torch.int32)
    assert torch.equal(union, torch.ones_like(union))
    # weighted (LPT) dealing: deterministic, a partition, and balanced to within the largest weight
    wts = torch.rand(43 * 75, generator=torch.Generator().manual_seed(0)) ** 4
    lpt = [parallel.TileShard(H, W, 4, r, weights=wts) for r in range(4)]
    assert torch.equal(sum(s.mask for s in lpt), torch.ones_like(total))
    loads = [float((wts.view(43, 75) * s.mask).sum()) for s in lpt]
    assert max(loads) - min(loads) <= float(wts.max()) + 1e-9
    assert torch.equal(lpt[1].mask, parallel.TileShard(H, W, 4, 1, weights=wts).mask)
    # single process: gather is the identity on owned pixels
    img = torch.rand(3, H, W)
    assert torch.equal(parallel.TileShard(H, W, 1, 0).gather(img), img)


# ---------------------------------------------------------------------------------------------- Gaussian sharding
def _gshard_worker(rank, world, port, q):
    """The two exchanges of a Gaussian-sharded frame (parallel.GaussianShard) over gloo: all-gather of equal row slices in
    place, reduce-scatter of equal row slices."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 12
        a, b = parallel.shard_range(P, world, rank)
        coll = parallel._Collectives(world, rank)
        rec = torch.full((P, 8), float("nan"))
        rec[a:b] = torch.arange(a, b, dtype=torch.float32)[:, None] * 10 + torch.arange(8, dtype=torch.float32)
        coll.all_gather_rows(rec)
        radii = torch.zeros(P, dtype=torch.int32)
        radii[a:b] = torch.arange(a, b, dtype=torch.int32) + 1
        coll.all_gather_rows(radii)
        partial = torch.full((P, 16), float(rank + 1))
        own = torch.empty((b - a, 16))
        coll.reduce_scatter_rows(partial, own)
        q.put((rank, rec.numpy(), radii.numpy(), own.numpy()))
    finally:
        dist.destroy_process_group()


def test_world2_gaussian_shard_exchanges():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gshard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(12, dtype=np.float32)[:, None] * 10 + np.arange(8, dtype=np.float32)
    for r in res:
        assert np.array_equal(r[1], want) and np.array_equal(r[2], np.arange(12) + 1)
        assert r[3].shape == (6, 16) and np.all(r[3] == 3.0)  # 1 + 2
    assert parallel.shard_range(12, 3, 2) == (8, 12)
    with pytest.raises(ValueError):
        parallel.shard_range(10, 4, 0)
    # single process: the exchanges degenerate to the identity / a copy
    c = parallel._Collectives(1, 0)
    x = torch.arange(6.0).view(3, 2)
    y = torch.empty(3, 2)
    c.all_gather_rows(x)
    c.reduce_scatter_rows(x, y)
    assert torch.equal(x, y)
