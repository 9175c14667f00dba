"""Multi-process (world_size 2 and 3, gloo, CPU) tests of the ray-sharding host logic:
range partition, feature broadcast and the all-gather of rendered tiles.  The CUDA render of
each shard is replaced by a deterministic function of the ray index, so what is tested is
exactly the plumbing that runs around car_render_forward on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cross_attention_renderer_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(total, b, R, rng, P=4):
    """Stand-in for model(input, z=z, ray_range=rng): full-size tensors, only the range is valid."""
    g = torch.arange(total, dtype=torch.float32)
    rgb = torch.full((b, 1, R, 3), float("nan"))
    vm = torch.full((b, R, 1), float("nan"))
    dep = torch.full((b, R, 1), float("nan"))
    lo, hi = rng
    rgb.view(total, 3)[lo:hi] = torch.stack([g, 2 * g, -g], -1)[lo:hi]
    vm.view(total, 1)[lo:hi] = (g % 2)[lo:hi, None]
    dep.view(total, 1)[lo:hi] = (g * 0.5)[lo:hi, None]
    return {"rgb": rgb, "valid_mask": vm, "depth_ray": dep}


def _worker(rank, world, port, b, R, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = b * R
        rng = sharding.ray_range(total, rank, world)
        out = _fake_render(total, b, R, rng)
        full = sharding.gather_tiles(out, total, rank, world)
        g = torch.arange(total, dtype=torch.float32)
        ok = torch.equal(full["rgb"].view(total, 3), torch.stack([g, 2 * g, -g], -1))
        ok &= torch.equal(full["valid_mask"].view(total), g % 2)
        ok &= torch.equal(full["depth_ray"].view(total), g * 0.5)
        # feature broadcast from rank 0
        z = [torch.full((2, 3), float(rank + 1)), torch.full((4,), float(10 * (rank + 1)))]
        sharding.broadcast_features(z, src=0)
        ok &= bool((z[0] == 1.0).all() and (z[1] == 10.0).all())
        q.put((rank, bool(ok), rng))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,b,R", [(2, 3, 37), (3, 2, 50), (2, 1, 5)])
def test_sharded_gather_gloo(world, b, R):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    ranges = sorted(r for _, _, r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == b * R
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0                                   # contiguous, no overlap, no gap


def test_ray_range_partition_properties():
    for total in (0, 1, 7, 64, 65537, 786432):
        for world in (1, 2, 3, 4, 8):
            rs = sharding.all_ranges(total, world)
            assert rs[0][0] == 0 and rs[-1][1] == total
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
            for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
                assert a1 == b0
    assert list(sharding.scenes_touched(10, 25, 10)) == [1, 2]
    assert list(sharding.scenes_touched(0, 0, 10)) == []


def _grad_worker(rank, world, port, average, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
        # rank-dependent gradients; the last layer has none on any rank (like the layers outside
        # the n_view=2 branch, which never receive a gradient)
        for i, p in enumerate(list(model.parameters())[:4]):
            p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
        n = sharding.average_gradients(model, average=average)
        tot = sum(r + 1 for r in range(world)) / (world if average else 1)
        ok = n == sum(p.numel() for p in list(model.parameters())[:4])
        for i, p in enumerate(list(model.parameters())[:4]):
            ok &= bool(torch.allclose(p.grad, torch.full_like(p, tot * (i + 1))))
        ok &= all(p.grad is None for p in list(model.parameters())[4:])
        q.put((rank, bool(ok), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,average", [(2, True), (3, False)])
def test_flat_gradient_allreduce_gloo(world, average):
    """One flat all_reduce replaces the reference's per-parameter loop (training.py:21-28)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, average, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


def _sync_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiment_scripts"))
    import _common as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # every rank starts from different weights
        model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        C.sync_model(model)
        flat = torch.cat([p.data.reshape(-1) for p in model.parameters()])
        ref = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(ref, flat)
        q.put((rank, all(torch.equal(r, ref[0]) for r in ref), None))
    finally:
        dist.destroy_process_group()


def test_sync_model_flat_broadcast_gloo():
    """Driver start-up: rank 0's weights reach every rank (train_realestate10k.py:60-62) in one broadcast."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


class _FakeModel:
    """model(input, z=z, ray_range=rng) stand-in: rgb depends on the ray index AND on the broadcast feature maps,
    so a rank that rendered with stale / un-broadcast maps is caught."""
    def __call__(self, inp, z=None, ray_range=None):
        b, _, R, _ = inp["query"]["uv"].shape
        total = b * R
        out = _fake_render(total, b, R, ray_range)
        tag = float(z[0].flatten()[0] + z[2].flatten()[-1])
        lo, hi = ray_range
        out["rgb"].view(total, 3)[lo:hi] += tag
        return out


def _pipeline_worker(rank, world, port, nscenes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, n, H, R = 1, 2, 8, 23
        scenes, want = [], []
        for k in range(nscenes):
            inp = {"context": {"rgb": torch.zeros(b, n, H, H, 3)}, "query": {"uv": torch.zeros(b, 1, R, 2)}}
            z = [torch.full(s, float(k + 1)) for s in sharding.feature_shapes(b, n, H, H)] if rank == 0 else None
            scenes.append((inp, z))
            g = torch.arange(b * R, dtype=torch.float32)
            want.append(torch.stack([g, 2 * g, -g], -1) + 2.0 * (k + 1))
        outs, nbytes = sharding.render_scenes_pipelined(_FakeModel(), scenes, src=0)
        ok = len(outs) == nscenes and nbytes == nscenes * sum(4 * torch.Size(s).numel() for s in sharding.feature_shapes(b, n, H, H))
        for o, w in zip(outs, want):
            ok &= torch.equal(o["rgb"].view(-1, 3), w)
        q.put((rank, bool(ok), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nscenes", [(2, 3), (3, 1)])
def test_pipelined_scene_broadcast_gloo(world, nscenes):
    """Strong-scaling path: features live on rank 0 only, are broadcast one scene ahead, every rank renders its
    ray range of every scene and the tiles are gathered."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, nscenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
