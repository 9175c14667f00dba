#!/usr/bin/env python
"""get_z (the reference's multi-view DPT-hybrid encoder, encoder.py) on the GPU: time per scene at 256x256 / 2 views,
memory order of its outputs, and the renderer running straight on them (forward(input) with z=None).
The encoder is plain torch (cuDNN / cuBLAS): this is a plumbing measurement, not a kernel claim."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cross_attention_renderer_b200 import packing, synthetic                      # noqa: E402
from cross_attention_renderer_b200.models import CrossAttentionRenderer            # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    b, H = int(os.environ.get("SCENES", "4")), 256
    torch.manual_seed(0)
    m = CrossAttentionRenderer(n_view=2, npoints=64, encoder="dpt_hybrid").to(dev).eval()
    m.pixel_val_to_cpu = False
    inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=1), dev)
    inp["context"]["rgb"] = torch.rand(b, 2, H, H, 3, device=dev) * 2 - 1
    out = {}
    for label, tf32 in (("tf32_convs_default", True), ("fp32_strict", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            for _ in range(3):
                z = m.get_z(inp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                z = m.get_z(inp)
            e1.record()
            torch.cuda.synchronize()
        out[label + "_ms_per_scene"] = round(e0.elapsed_time(e1) / 5 / b, 3)
    torch.backends.cudnn.allow_tf32 = True
    out["outputs_nhwc_zero_copy"] = all(packing.nhwc_view(t) is not None for t in z)
    with torch.no_grad():
        for _ in range(2):
            res = m(inp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            res = m(inp)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
    out["forward_with_encoder_ms_per_scene"] = round(dt * 1e3 / b, 2)
    out["rays_per_s_with_encoder"] = round(b * H * H / dt, 1)
    out["rgb_finite"] = bool(torch.isfinite(res["rgb"]).all())
    out["scenes"] = b
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
