#!/usr/bin/env python
"""Training step at BASELINE config 5 sizes (12 scenes x 192 rays per GPU, 64 samples, 256x256 maps):
train-mode forward + backward through the CUDA path + the optimiser step (gradient all-reduce across
ranks, clip_grad_norm_(1), Adam) - reference training.py:86-136.  Encoder excluded (z given and requiring
grad, so the feature-map scatter is included).  Prints one JSON line on rank 0.

    python scripts/train_bench.py                                            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 scripts/train_bench.py   # DDP
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cross_attention_renderer_b200 import _lib, synthetic                      # noqa: E402
from cross_attention_renderer_b200.models import CrossAttentionRenderer        # noqa: E402
from cross_attention_renderer_b200.optim import FlatAdam                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=12)
    ap.add_argument("--rays", type=int, default=192)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-feature-grads", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="A/B: per-parameter all-reduce + clip + torch.optim.Adam")
    ap.add_argument("--with-encoder", action="store_true",
                    help="C5 as the reference trains it: get_z (multi-view DPT-hybrid, encoder.py) inside the step, its "
                         "parameters in the optimiser")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_simt"],
                    help="fp32: per-sample GEMMs of forward and backward on tcgen05 (hi + lo bf16); fp32_simt: exact fp32")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    inp = synthetic.to_device(synthetic.make_inputs(a.scenes, a.size, a.size, seed=rank, rays=a.rays), dev)
    z = [t.to(dev).requires_grad_(not a.no_feature_grads) for t in synthetic.make_features(a.scenes, a.size, seed=rank)]
    m = CrossAttentionRenderer(n_view=2, npoints=a.samples, precision=a.precision,
                               encoder="dpt_hybrid" if a.with_encoder else None).to(dev)
    if a.with_encoder:
        inp["context"]["rgb"] = torch.rand(a.scenes, 2, a.size, a.size, 3, device=dev) * 2 - 1
    m.load_state_dict(synthetic.make_state_dict(seed=0), strict=False)
    m.H = m.W = a.size
    m.train()
    m.pixel_val_to_cpu = False
    R = inp["query"]["uv"].shape[2]
    target = torch.rand(a.scenes, 1, R, 3, device=dev) * 2 - 1
    if a.torch_adam:
        opt = torch.optim.Adam(m.parameters(), lr=5e-5, betas=(0.99, 0.999))
    else:
        opt = FlatAdam(m.parameters(), lr=5e-5, betas=(0.99, 0.999))

    marks = []

    def mark():
        if marks is not None and len(marks) < 4:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    def step():
        opt.zero_grad()
        for t in z:
            t.grad = None
        mark()
        out = m(inp) if a.with_encoder else m(inp, z=z)
        loss = (out["rgb"] - target).abs().mean()          # image_loss (loss_functions.py:74-80)
        mark()
        loss.backward()
        mark()
        if a.torch_adam:
            if world > 1:                                   # training.py:21-28: one all_reduce per parameter
                for p in m.parameters():
                    if p.grad is not None:
                        dist.all_reduce(p.grad.data, op=dist.ReduceOp.SUM)
                        p.grad.data /= float(world)
            torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1.0)
            opt.step()
        else:
            opt.step(max_grad_norm=1.0)
        mark()
        return loss.detach()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    sync()
    marks.clear()                  # one step split into forward / backward / all-reduce + optimiser (device time)
    step()
    sync()
    split = {k: round(marks[i].elapsed_time(marks[i + 1]), 3) for i, k in enumerate(("forward_ms", "backward_ms", "allreduce_optim_ms"))}
    marks = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # one profiled step: kernel time per stage (this rank)
    import ctypes as C
    lib.car_profile_begin()
    step()
    n = len(_lib.STAGES)
    tms, cnt = (C.c_float * n)(), (C.c_int * n)()
    lib.car_profile_end(tms, cnt, n)
    stages = {s: round(float(tms[i]), 3) for i, s in enumerate(_lib.STAGES) if cnt[i]}
    rays = a.scenes * R * world
    if rank == 0:
        print(json.dumps({"metric": "train_step_rays_per_s (renderer fwd+bwd + optimiser step, encoder "
                                    + ("INCLUDED: get_z fwd+bwd, 123.6 M parameters in the optimiser)" if a.with_encoder else "excluded)"),
                          "value": round(rays / ms * 1e3, 1), "unit": "rays/s", "ms_per_step": round(ms, 3), "n_gpus": world,
                          "config": {"scenes_per_gpu": a.scenes, "rays_per_scene": R, "samples": a.samples, "size": a.size,
                                     "feature_grads": not a.no_feature_grads, "precision": a.precision,
                                     "optimizer": "torch.optim.Adam + per-parameter all_reduce + clip_grad_norm_" if a.torch_adam
                                     else "FlatAdam: one all_reduce + clip scalar + car_adam_step"},
                          "step_split": split, "kernel_ms_by_stage": stages, "loss": float(loss)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
