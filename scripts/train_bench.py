#!/usr/bin/env python
"""Renderer-only training step at BASELINE config 5 sizes (12 scenes x 192 rays, 64 samples,
256x256 maps): train-mode forward + backward through the CUDA path, encoder excluded (z given
and requiring grad, so the feature-map scatter is included).  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cross_attention_renderer_b200 import _lib, synthetic                      # noqa: E402
from cross_attention_renderer_b200.models import CrossAttentionRenderer        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=12)
    ap.add_argument("--rays", type=int, default=192)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-feature-grads", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    lib = _lib.load()
    inp = synthetic.to_device(synthetic.make_inputs(a.scenes, a.size, a.size, seed=0, rays=a.rays), dev)
    z = [t.to(dev).requires_grad_(not a.no_feature_grads) for t in synthetic.make_features(a.scenes, a.size, seed=0)]
    m = CrossAttentionRenderer(n_view=2, npoints=a.samples).to(dev)
    m.load_state_dict(synthetic.make_state_dict(seed=0), strict=False)
    m.H = m.W = a.size
    m.train()
    m.pixel_val_to_cpu = False
    R = inp["query"]["uv"].shape[2]
    target = torch.rand(a.scenes, 1, R, 3, device=dev) * 2 - 1

    def step(profile=False):
        for p in m.parameters():
            p.grad = None
        for t in z:
            t.grad = None
        out = m(inp, z=z)
        loss = (out["rgb"] - target).abs().mean()          # image_loss (loss_functions.py:74-80)
        loss.backward()
        return float(loss.detach())

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # one profiled step: kernel time per stage
    import ctypes as C
    lib.car_profile_begin()
    step()
    n = len(_lib.STAGES)
    tms, cnt = (C.c_float * n)(), (C.c_int * n)()
    lib.car_profile_end(tms, cnt, n)
    stages = {s: round(float(tms[i]), 3) for i, s in enumerate(_lib.STAGES) if cnt[i]}
    rays = a.scenes * R
    print(json.dumps({"metric": "train_step_rays_per_s (renderer fwd+bwd, encoder excluded)",
                      "value": round(rays / ms * 1e3, 1), "unit": "rays/s", "ms_per_step": round(ms, 3),
                      "config": {"scenes": a.scenes, "rays_per_scene": R, "samples": a.samples, "size": a.size,
                                 "feature_grads": not a.no_feature_grads},
                      "kernel_ms_by_stage": stages, "loss": loss}))


if __name__ == "__main__":
    main()
