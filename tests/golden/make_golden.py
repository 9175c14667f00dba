#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  The reference is imported from where it lies; nothing is
copied.  Third-party modules the reference imports at module scope but that
are absent here (timm, matplotlib, midas internals -> timm) are replaced by
empty stub modules, and ``Tensor.cuda`` is made a no-op because the reference
hard-codes ``.cuda()`` in geometry.py:320,398.  None of the stubs is on the
arithmetic path of ``CrossAttentionRenderer.forward(input, z=z)``.

Outputs (committed, small): tests/golden/<case>.npz with the inputs'
generator arguments, every ``out_dict`` entry, and intermediates captured with
forward hooks / a grid_sample tap.

    python tests/golden/make_golden.py            # writes all cases
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = os.environ.get("CAR_REFERENCE_DIR", "/root/reference")


def import_reference():
    sys.dont_write_bytecode = True          # reference tree is read-only
    for name in ("matplotlib", "matplotlib.colors", "timm", "timm.models",
                 "timm.models.layers", "timm.models.layers.std_conv",
                 "midas", "midas.dpt_depth", "midas.midas_net",
                 "midas.midas_net_custom"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    timm = sys.modules["timm"]
    timm.models = sys.modules["timm.models"]
    timm.models.layers = sys.modules["timm.models.layers"]
    timm.models.layers.std_conv = sys.modules["timm.models.layers.std_conv"]

    class _StdConv2dSame(torch.nn.Conv2d):
        def __init__(self, cin, cout, kernel_size, stride, bias):
            super().__init__(cin, cout, kernel_size, stride=stride, bias=bias)
    timm.models.layers.std_conv.StdConv2dSame = _StdConv2dSame

    class _NS(torch.nn.Module):
        def __init__(self):
            super().__init__()

    class _DPT(torch.nn.Module):           # encoder placeholder (not on the path)
        def __init__(self, **kw):
            super().__init__()
            self.pretrained = _NS()
            self.pretrained.model = _NS()
            self.pretrained.model.patch_embed = _NS()
            self.pretrained.model.patch_embed.backbone = _NS()
            self.pretrained.model.patch_embed.backbone.stem = _NS()
    sys.modules["midas.dpt_depth"].DPTDepthModel = _DPT
    midas = sys.modules["midas"]
    midas.dpt_depth = sys.modules["midas.dpt_depth"]
    midas.midas_net = sys.modules["midas.midas_net"]
    midas.midas_net_custom = sys.modules["midas.midas_net_custom"]
    torch.Tensor.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models as ref_models            # noqa: E402  (reference models.py)
    return ref_models


CASES = {
    # name: dict(b, H, Ht, P, seed, mode, peaky, rays)
    # ``tap_rays``: intermediates are stored for that many evenly spaced rays only
    # (out_dict entries and sampling grids are always stored in full).
    "tiny_default": dict(b=1, H=32, Ht=8, P=8, seed=0, mode="default", peaky=False, rays=None, tap_rays=4, keep_feat=True),
    "tiny_peaky": dict(b=2, H=32, Ht=8, P=16, seed=1, mode="default", peaky=True, rays=None, tap_rays=2, keep_feat=True),
    "tiny_mixed": dict(b=3, H=32, Ht=8, P=8, seed=2, mode="mixed", peaky=False, rays=None, tap_rays=3, keep_feat=True),
    "c1_sparse": dict(b=1, H=64, Ht=64, P=32, seed=3, mode="default", peaky=False, rays=96, tap_rays=1, keep_feat=False),
    "p64_sparse": dict(b=1, H=64, Ht=64, P=64, seed=4, mode="default", peaky=True, rays=48, tap_rays=1, keep_feat=False),
}


def run_case(ref_models, cfg):
    sys.path.insert(0, REPO)
    from cross_attention_renderer_b200 import synthetic
    import torch.nn.functional as F

    inp = synthetic.make_inputs(cfg["b"], cfg["H"], cfg["Ht"], seed=cfg["seed"],
                                mode=cfg["mode"], rays=cfg["rays"])
    z = synthetic.make_features(cfg["b"], cfg["H"], seed=cfg["seed"])
    sd = synthetic.make_state_dict(seed=cfg["seed"], peaky=cfg["peaky"])

    torch.manual_seed(0)
    m = ref_models.CrossAttentionRenderer(model="midas_vit", n_view=2, npoints=cfg["P"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("encoder.") for k in missing), missing
    m.eval()
    m.H = m.W = cfg["H"]

    taps = {}
    hooks = []

    def hook(name):
        def fn(mod, args, out):
            taps.setdefault(name, []).append(out.detach().clone())
        return fn
    for lname in ("query_encode_latent_2", "latent_value", "key_map_2", "query_embed_2",
                  "encode_latent", "query_repeat_embed_2"):
        hooks.append(getattr(m, lname).register_forward_hook(hook(lname)))

    grids = []
    orig_gs = F.grid_sample

    def gs(inp_, grid, **kw):
        out = orig_gs(inp_, grid, **kw)
        grids.append((kw.get("padding_mode"), grid.detach().clone(), out.detach().clone()))
        return out
    ref_models.F.grid_sample = gs
    try:
        with torch.no_grad():
            out = m(inp, z=z)
    finally:
        ref_models.F.grid_sample = orig_gs
        for h in hooks:
            h.remove()

    rec = {}
    for k in ("rgb", "valid_mask", "depth_ray", "at_wt", "at_wt_max", "pixel_val", "coords"):
        rec["out_" + k] = out[k].detach().cpu().numpy()
    # grid_sample calls: 3 border (primary), 3 zeros (cross-view) — models.py:278,317
    border = [g for g in grids if g[0] == "border"]
    zeros = [g for g in grids if g[0] == "zeros"]
    assert len(border) == 3 and len(zeros) == 3
    rec["grid_primary"] = border[0][1].numpy()
    rec["grid_cross"] = zeros[0][1].numpy()
    R = rec["grid_primary"].shape[1]
    ridx = torch.linspace(0, R - 1, cfg["tap_rays"]).round().long()
    rec["tap_ray_index"] = ridx.numpy()
    if cfg["keep_feat"]:
        rec["feat_primary"] = torch.cat([g[2] for g in border], 1)[:, :, ridx].numpy()   # (bn,576,r,P)
        rec["feat_cross"] = torch.cat([g[2] for g in zeros], 1)[:, :, ridx].numpy()
    # query_encode_latent_2 is called 4×: (1_enc_1, 1_enc_2, 2_enc_1, 2_enc_2) models.py:333-341
    enc = taps["query_encode_latent_2"]
    assert len(enc) == 4
    rec["enc"] = torch.stack(enc, 0)[:, :, :, ridx].numpy()                # (4,b,288,r,P)
    rec["value"] = taps["latent_value"][0][:, :, ridx].numpy()             # (bn,288,r,P)
    rec["key"] = taps["key_map_2"][0][:, :, ridx].numpy()                  # (bn,128,r,P)
    rec["q1"] = taps["query_embed_2"][0][:, :, ridx].numpy()
    rec["g"] = taps["encode_latent"][0].numpy()                            # (bn,128,R)
    rec["q2"] = taps["query_repeat_embed_2"][0][:, :, ridx].numpy()
    rec["cfg"] = np.array(repr(cfg))
    return rec


def main():
    ref_models = import_reference()
    torch.set_num_threads(os.cpu_count())
    outdir = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        rec = run_case(ref_models, cfg)
        # keep fixtures small: big intermediates stored as float16-free fp32 but sub-sampled
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape")},
              "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
