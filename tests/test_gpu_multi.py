"""Multi-GPU (NCCL) tests of the ray-sharded path on real hardware: one process per GPU, features present on
rank 0 only and broadcast, every rank renders its ray range through car_render_forward, tiles all-gathered -
sharded == unsharded, bit for bit.  Needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cross_attention_renderer_b200 import sharding, synthetic

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    from cross_attention_renderer_b200.models import CrossAttentionRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        b, H, Ht, P = 2, 64, 24, 64
        sd = synthetic.make_state_dict(seed=81)
        model = CrossAttentionRenderer(n_view=2, npoints=P, precision="fp32").to(dev).eval()
        model.load_state_dict(sd, strict=False)
        model.H = model.W = H
        scenes = []
        for k in range(3):
            inp = synthetic.to_device(synthetic.make_inputs(b, H, Ht, seed=81 + k, mode="mixed"), dev)
            z = [t.to(dev) for t in synthetic.make_features(b, H, seed=81 + k)] if rank == 0 else None
            scenes.append((inp, z))
        with torch.no_grad():
            outs, nbytes = sharding.render_scenes_pipelined(model, scenes, src=0, device=dev)
            ok = True
            for k, (inp, _) in enumerate(scenes):
                zfull = [t.to(dev) for t in synthetic.make_features(b, H, seed=81 + k)]
                ref = model(inp, z=zfull)                             # unsharded, on this rank
                for key in ("rgb", "valid_mask", "depth_ray"):
                    ok &= bool(torch.equal(outs[k][key], ref[key]))
                lo, hi = outs[k]["ray_range"]
                ok &= (lo, hi) == sharding.ray_range(b * Ht * Ht, rank, world)
            # flat gradient all-reduce over NCCL (training.py:21-28 replacement)
            lin = torch.nn.Linear(4, 3).to(dev)
            for p in lin.parameters():
                p.grad = torch.full_like(p, float(rank + 1))
            sharding.average_gradients(lin)
            ok &= all(bool(torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2.0))) for p in lin.parameters())
        torch.cuda.synchronize()
        q.put((rank, bool(ok), int(nbytes)))
    finally:
        dist.destroy_process_group()


def test_sharded_render_equals_unsharded_nccl():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(nb > 0 for _, _, nb in res)
