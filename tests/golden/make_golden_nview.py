#!/usr/bin/env python
"""Golden vectors of the reference's OTHER forward branches (n_view = 1, n_view = 3), produced by the
UNMODIFIED reference on CPU exactly like make_golden.py (same import stubs).  They pin the oracle
restatements ``car_oracle.render_single_view`` / ``render_three_views``; the CUDA path for these
branches is a later row.

    python tests/golden/make_golden_nview.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REPO, import_reference                      # noqa: E402

CASES = {
    "nview1_default": dict(n_view=1, b=2, H=32, Ht=8, P=16, seed=5, mode="default", peaky=True),
    "nview1_mixed": dict(n_view=1, b=3, H=32, Ht=8, P=8, seed=6, mode="mixed", peaky=False),
    "nview3_default": dict(n_view=3, b=2, H=32, Ht=8, P=8, seed=7, mode="default", peaky=True),
    "nview3_mixed": dict(n_view=3, b=3, H=32, Ht=6, P=8, seed=8, mode="mixed", peaky=False),
    "nview2_nosample": dict(n_view=2, b=2, H=32, Ht=8, P=16, seed=9, mode="mixed", peaky=True, no_sample=True),
    "nview2_nolatentconcat": dict(n_view=2, b=2, H=32, Ht=8, P=8, seed=10, mode="default", peaky=True,
                                  no_latent_concat=True),
}


def run_case(ref_models, cfg):
    sys.path.insert(0, REPO)
    from cross_attention_renderer_b200 import synthetic
    from cross_attention_renderer_b200.params import renderer_param_shapes
    nv = cfg["n_view"]
    inp = synthetic.make_inputs(cfg["b"], cfg["H"], cfg["Ht"], seed=cfg["seed"], mode=cfg["mode"], n_ctx=nv)
    z = synthetic.make_features(cfg["b"], cfg["H"], seed=cfg["seed"], n_view=nv)
    nlc, nos = cfg.get("no_latent_concat", False), cfg.get("no_sample", False)
    sd = synthetic.make_state_dict(seed=cfg["seed"], peaky=cfg["peaky"], n_view=nv, no_latent_concat=nlc)
    torch.manual_seed(0)
    m = ref_models.CrossAttentionRenderer(model="midas_vit", n_view=nv, npoints=cfg["P"], no_sample=nos,
                                          no_latent_concat=nlc)
    # the checkpoint ABI restated in params.py must be the reference module's own parameter list
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("encoder.")}
    assert ref_shapes == {k: tuple(v) for k, v in renderer_param_shapes(nv, no_latent_concat=nlc).items()}, "params.py != reference"
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing)
    m.eval()
    m.H = m.W = cfg["H"]
    with torch.no_grad():
        out = m(inp, z=z)
    rec = {"out_" + k: out[k].detach().cpu().numpy()
           for k in ("rgb", "valid_mask", "depth_ray", "at_wt", "at_wt_max", "pixel_val", "coords")}
    rec["cfg"] = np.array(repr(cfg))
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
        print(name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape") and v.shape}, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
