"""Pins the oracle's gradients (oracle.car_oracle.render_grad) to golden gradients produced by
autograd through the unmodified reference (tests/golden/make_golden_grad.py)."""
import ast
import os

import numpy as np
import pytest
import torch

from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.params import HOT_PATH_PARAMS
from oracle import car_oracle as oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["grad_tiny_mixed", "grad_tiny_peaky"]


def load_grad_case(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(d["cfg"]))
    inp = synthetic.make_inputs(cfg["b"], cfg["H"], cfg["Ht"], seed=cfg["seed"], mode=cfg["mode"])
    z = synthetic.make_features(cfg["b"], cfg["H"], seed=cfg["seed"])
    sd = synthetic.make_state_dict(seed=cfg["seed"], peaky=cfg["peaky"])
    R = cfg["Ht"] * cfg["Ht"]
    g = torch.Generator().manual_seed(1000 + cfg["seed"])
    g_rgb = torch.randn(cfg["b"], 1, R, 3, generator=g)
    g_depth = torch.randn(cfg["b"], R, 1, generator=g) * 0.25
    return d, cfg, inp, z, sd, g_rgb, g_depth


def check_against_golden(d, name, grad, tol, entry_tol=None, tail=True):
    """Compare a gradient tensor with the stored strided sub-sample + norm.  ``entry_tol`` (default: ``tol``)
    bounds the worst single entry; the norm is held to ``tol`` and so are 99.5 % of the sampled entries
    (``tail``; without it only the median entry, at 5e-3 of the rms)."""
    ref = torch.from_numpy(d["g:" + name])
    stride = int(d["s:" + name])
    nrm = float(d["n:" + name])
    got = grad.detach().cpu().reshape(-1)[::stride]
    scale = max(nrm / max(grad.numel(), 1) ** 0.5, 1e-12)        # rms of the reference gradient
    rel = (got - ref).abs() / scale
    err = float(rel.max())
    q = float(torch.quantile(rel.double(), 0.995)) if rel.numel() > 1 else err
    nerr = abs(float(grad.double().norm()) - nrm) / max(nrm, 1e-12)
    if not tail:
        q = float(rel.median()) * (tol / 5e-3)          # median <= 5e-3 of the rms
    assert err < (entry_tol or tol) and q < tol and nerr < tol, (name, err, q, nerr, int((rel >= tol).sum()), rel.numel())


@pytest.mark.parametrize("case", CASES)
def test_oracle_gradients_match_reference(case):
    d, cfg, inp, z, sd, g_rgb, g_depth = load_grad_case(case)
    out, grads, gz = oracle.render_grad(sd, inp, z, cfg["H"], cfg["H"], cfg["P"], g_rgb, g_depth)
    assert np.abs(out["rgb"].detach().numpy() - d["out_rgb"]).max() < 2e-5
    assert np.abs(out["depth_ray"].detach().numpy() - d["out_depth_ray"]).max() < 2e-4
    # max-abs error relative to the gradient's rms: matmul order differs, nothing else
    for name in HOT_PATH_PARAMS:
        check_against_golden(d, name, grads[name], 2e-3)
    for i in range(3):
        check_against_golden(d, f"z{i}", gz[i], 2e-3)
    assert list(d["untouched"]) == []
