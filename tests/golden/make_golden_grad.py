#!/usr/bin/env python
"""Golden GRADIENTS: autograd through the UNMODIFIED reference on CPU.

Same stubbing as make_golden.py (imported from there).  For each case the scalar

    L = sum(out['rgb'] * G_rgb) + sum(out['depth_ray'] * G_depth)

(fixed seeded cotangents) is back-propagated with ``L.backward()`` exactly like the
reference's training step does for its own loss (training.py:125), and the gradients of
every hot-path parameter and of the three feature maps are stored: small tensors in full,
big ones as a fixed strided sub-sample plus their L2 norm.

    python tests/golden/make_golden_grad.py     # writes tests/golden/grad_*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
from make_golden import import_reference          # noqa: E402

MAX_KEEP = 16384

CASES = {
    "grad_tiny_mixed": dict(b=3, H=32, Ht=8, P=8, seed=5, mode="mixed", peaky=False),
    "grad_tiny_peaky": dict(b=1, H=32, Ht=8, P=16, seed=6, mode="default", peaky=True),
}


def cotangents(cfg, R):
    g = torch.Generator().manual_seed(1000 + cfg["seed"])
    return (torch.randn(cfg["b"], 1, R, 3, generator=g),
            torch.randn(cfg["b"], R, 1, generator=g) * 0.25)


def subsample(t):
    flat = t.detach().reshape(-1)
    stride = max(1, -(-flat.numel() // MAX_KEEP))
    return flat[::stride].numpy().copy(), stride, float(flat.double().norm())


def run_case(ref_models, cfg):
    from cross_attention_renderer_b200 import synthetic
    from cross_attention_renderer_b200.params import HOT_PATH_PARAMS
    inp = synthetic.make_inputs(cfg["b"], cfg["H"], cfg["Ht"], seed=cfg["seed"], mode=cfg["mode"])
    z = [t.clone().requires_grad_(True) for t in synthetic.make_features(cfg["b"], cfg["H"], seed=cfg["seed"])]
    sd = synthetic.make_state_dict(seed=cfg["seed"], peaky=cfg["peaky"])
    torch.manual_seed(0)
    m = ref_models.CrossAttentionRenderer(model="midas_vit", n_view=2, npoints=cfg["P"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    m.train()
    m.H = m.W = cfg["H"]
    out = m(inp, z=z)
    R = out["rgb"].shape[2]
    g_rgb, g_depth = cotangents(cfg, R)
    loss = (out["rgb"] * g_rgb).sum() + (out["depth_ray"] * g_depth).sum()
    loss.backward()
    rec = {"cfg": np.array(repr(cfg)), "loss": np.array(float(loss)),
           "out_rgb": out["rgb"].detach().numpy(), "out_depth_ray": out["depth_ray"].detach().numpy()}
    params = dict(m.named_parameters())
    for name in HOT_PATH_PARAMS:
        g = params[name].grad
        assert g is not None, name
        v, stride, nrm = subsample(g)
        rec["g:" + name] = v
        rec["s:" + name] = np.array(stride)
        rec["n:" + name] = np.array(nrm)
    # parameters outside the n_view=2 branch must not receive a gradient
    rec["untouched"] = np.array([n for n, p in params.items()
                                 if not n.startswith("encoder.") and n not in HOT_PATH_PARAMS and p.grad is not None])
    for i, t in enumerate(z):
        v, stride, nrm = subsample(t.grad)
        rec[f"g:z{i}"] = v
        rec[f"s:z{i}"] = np.array(stride)
        rec[f"n:z{i}"] = np.array(nrm)
    return rec


def main():
    ref_models = import_reference()
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        rec = run_case(ref_models, cfg)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "loss %.6f" % float(rec["loss"]), "untouched-with-grad:", list(rec["untouched"]),
              "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
