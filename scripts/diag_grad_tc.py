#!/usr/bin/env python
"""Per-tensor gradient error statistics of the training path against the oracle's autograd, for the three GEMM
configurations: exact fp32 ("simt"), exact forward + tensor-core backward ("hybrid"), tensor cores for both ("tc").
Separates the GEMM arithmetic error from the ReLU-flip outliers (tests/test_gpu_backward.py: ENTRY_TOL / L2_TOL)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
from cross_attention_renderer_b200 import synthetic                               # noqa: E402
from cross_attention_renderer_b200.params import HOT_PATH_PARAMS                  # noqa: E402
from oracle import car_oracle as orc                                             # noqa: E402
import test_gpu_backward as T                                                    # noqa: E402

CASES = [(2, 64, 16, 64, "default", False, True, 21), (3, 32, 8, 8, "mixed", False, True, 21),
         (1, 32, 12, 32, "mixed", True, True, 21), (2, 64, 16, 64, "default", False, True, 31)]


def stats(got, ref):
    d = (got - ref).double().abs().reshape(-1)
    r = ref.double().abs().reshape(-1)
    rms = float(ref.double().norm()) / max(ref.numel(), 1) ** 0.5
    if rms == 0:
        return None
    return {"l2": float(d.norm() / ref.double().norm()), "max_rms": float(d.max() / rms),
            "max_scaled": float((d / torch.clamp(r, min=rms)).max()), "median_rms": float(d.median() / rms),
            "frac_gt_1e-3": float((d > 1e-3 * rms).double().mean()), "kurt": float((r ** 4).mean() / (r ** 2).mean() ** 2)}


def main():
    for (b, H, Ht, P, mode, peaky, depth, seed) in CASES:
        inp = synthetic.make_inputs(b, H, Ht, seed=seed, mode=mode)
        z = synthetic.make_features(b, H, seed=seed)
        sd = synthetic.make_state_dict(seed=seed, peaky=peaky)
        g = torch.Generator().manual_seed(5)
        g_rgb = torch.randn(b, 1, Ht * Ht, 3, generator=g)
        g_depth = torch.randn(b, Ht * Ht, 1, generator=g) * 0.25 if depth else None
        cams = orc.prepare_cameras(inp)
        _, ref, ref_z = orc.render_grad(sd, inp, z, H, H, P, g_rgb, g_depth, cams=cams)
        for label, fwd, bwd in (("simt", "fp32_simt", None), ("hybrid", "fp32_simt", "fp32"), ("tc", "fp32", None)):
            m = T.make_model(sd, P, H, fwd)
            if bwd:
                m.backward_precision = bwd
            _, grads, gz = T.cuda_grads(m, inp, z, cams, P, g_rgb, g_depth)
            rows = {n: stats(grads[n], ref[n]) for n in HOT_PATH_PARAMS}
            rows.update({f"z{i}": stats(gz[i], ref_z[i]) for i in range(3)})
            rows = {k: v for k, v in rows.items() if v}
            worst = sorted(rows.items(), key=lambda kv: -kv[1]["max_scaled"])[:4]
            agg = {k: max(v[k] for v in rows.values()) for k in ("l2", "max_rms", "max_scaled", "median_rms", "frac_gt_1e-3")}
            print(json.dumps({"case": [b, H, Ht, P, mode, peaky, depth, seed], "mode": label, "worst_over_tensors": {k: float(f"{v:.3g}") for k, v in agg.items()},
                              "worst4": {k: {kk: float(f"{vv:.3g}") for kk, vv in v.items()} for k, v in worst}}), flush=True)


if __name__ == "__main__":
    main()
