#!/usr/bin/env python
"""A/B of the bilinear tap fetch (DESIGN.md "tap fetch A/B"): the fused kernel's LDG
producers against TMA ``tile::gather4`` staging, on one K-stage-shaped workload (64 sample rows x 32 fp32
channels per stage, four 128-byte taps per row), csrc/car_tap_fetch_ab.cu in libcar_b200_test.so.

    python scripts/tap_fetch_ab.py            # every configuration, each in its own process, one JSON line each

Rows are samples along random line segments of an h x w NHWC level (C = 256), like the epipolar samples of a
ray; the level is 16 MB (L2-resident, the common case of the renderer) or 1 GB (``--big``: HBM-resident)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cross_attention_renderer_b200 import _lib                                   # noqa: E402


def make_case(h, w, Cc, rays, P, seed, dev):
    g = torch.Generator().manual_seed(seed)
    fmap = torch.randn(h * w, Cc, generator=g).to(dev)
    a = torch.rand(rays, 2, generator=g) * torch.tensor([w - 1.0, h - 1.0])
    b = torch.rand(rays, 2, generator=g) * torch.tensor([w - 1.0, h - 1.0])
    t = torch.linspace(0, 1, P)[None, :, None]
    pts = (a[:, None] * (1 - t) + b[:, None] * t).reshape(-1, 2)
    x0, y0 = pts[:, 0].floor().long().clamp(0, w - 2), pts[:, 1].floor().long().clamp(0, h - 2)
    fx, fy = (pts[:, 0] - x0).clamp(0, 1), (pts[:, 1] - y0).clamp(0, 1)
    taps = torch.stack([y0 * w + x0, y0 * w + x0 + 1, (y0 + 1) * w + x0, (y0 + 1) * w + x0 + 1], 1).int()
    wts = torch.stack([(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy], 1).float()
    return fmap, taps.contiguous().to(dev), wts.contiguous().to(dev)


def run_one(a):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load_test()
    h = w = 1024 if a.big else 128
    fmap, taps, wts = make_case(h, w, 256, a.rays, 64, 0, dev)
    n = taps.shape[0]
    out = torch.zeros(n, 256, dtype=torch.bfloat16, device=dev)
    ms = C.c_float()
    rc = lib.car_tap_fetch_ab(fmap.data_ptr(), h * w, 256, taps.data_ptr(), wts.data_ptr(), out.data_ptr(), n,
                              a.variant, a.box_rows, a.nslot, a.ctas_per_sm, a.iters, C.byref(ms),
                              torch.cuda.current_stream().cuda_stream, a.issue_warps)
    res = {"variant": "ldg" if a.variant == 0 else "tma_gather4", "box_rows": a.box_rows if a.variant else None,
           "nslot": a.nslot if a.variant else None, "issue_warps": a.issue_warps if a.variant else None, "ctas_per_sm": a.ctas_per_sm, "map_mb": h * w * 256 * 4 / 2**20,
           "rows": n, "rc": rc}
    if rc == 0:
        k = 8192
        ref = (fmap[taps[:k].long()] * wts[:k, :, None]).sum(1)       # (k, 4, C) -> (k, C); fma order differs: compare in bf16 ulps
        err = float((out[:k].float() - ref).abs().max() / ref.abs().max())
        tap_bytes = n * 4 * 256 * 4
        res.update(ms=round(ms.value, 4), tap_gb_per_s=round(tap_bytes / ms.value / 1e6, 1), max_rel_err=err,
                   ok=err < 1e-2)
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--box_rows", type=int, default=1)
    ap.add_argument("--nslot", type=int, default=4)
    ap.add_argument("--ctas_per_sm", type=int, default=0)
    ap.add_argument("--issue_warps", type=int, default=1)
    ap.add_argument("--rays", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    if a.one:
        return run_one(a)
    cfgs = [dict(variant=0, ctas_per_sm=c) for c in (1, 2, 3, 4)]
    cfgs += [dict(variant=1, box_rows=1, nslot=ns, issue_warps=iw) for ns in (2, 6) for iw in (1, 2)]
    cfgs += [dict(variant=1, box_rows=4, nslot=4)]        # rejected by the hardware (gather4 wants a one-row box)
    for big in (False, True):
        for c in cfgs:
            cmd = [sys.executable, os.path.abspath(__file__), "--one", "--rays", str(a.rays), "--iters", str(a.iters)]
            cmd += sum(([f"--{k}", str(v)] for k, v in c.items()), []) + (["--big"] if big else [])
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                print(line[-1] if line else json.dumps({"cfg": c, "big": big, "failed": (r.stderr or r.stdout)[-300:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"cfg": c, "big": big, "failed": "timeout"}), flush=True)


if __name__ == "__main__":
    main()
