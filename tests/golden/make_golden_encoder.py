#!/usr/bin/env python
"""Golden vectors for the image encoder, produced by the UNMODIFIED reference wrapper:
``midas.dpt_depth.DPTDepthModel`` (reference midas/dpt_depth.py, midas/blocks.py, midas/vit.py - read-outs,
reassemble stacks, scratch convolutions, fusion blocks, ``forward_vit`` and the multi-view ``forward_flex``)
imported from /root/reference, built around the ViT-hybrid of ``cross_attention_renderer_b200/encoder.py``
(the part the reference takes from timm 0.5.4, which is not installed here: ``timm`` is an empty stub and
``vit_models.vit_base_resnet50_384`` returns our restatement).

    python tests/golden/make_golden_encoder.py      # writes tests/golden/encoder_golden.npz (build container only)

Stored: the seed, a strided sub-sample of both output maps and their norms (the maps are 12 MB)."""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("CAR_REFERENCE_DIR", "/root/reference")
SEED, STRIDE = 1234, 997


@contextlib.contextmanager
def reference_midas():
    """The reference's real ``midas`` package with ``timm`` / ``vit_models`` stubbed; restores sys.modules after."""
    from cross_attention_renderer_b200.encoder import VisionTransformerMultiView
    saved = {k: v for k, v in sys.modules.items() if k == "midas" or k.startswith("midas.") or k in ("timm", "vit_models")}
    for k in saved:
        del sys.modules[k]
    sys.modules["timm"] = types.ModuleType("timm")
    vm = types.ModuleType("vit_models")
    vm.vit_base_resnet50_384 = lambda pretrained=False, **kw: VisionTransformerMultiView()
    sys.modules["vit_models"] = vm
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        import midas.dpt_depth as dpt_depth
        yield dpt_depth
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "midas" or k.startswith("midas.") or k in ("timm", "vit_models")]:
            del sys.modules[k]
        sys.modules.update(saved)


def inputs():
    g = torch.Generator().manual_seed(SEED + 1)
    return torch.randn(2, 3, 256, 256, generator=g), torch.randn(2, 16, generator=g) * 0.5


def build_ours():
    from cross_attention_renderer_b200.encoder import DPTHybridEncoder
    torch.manual_seed(SEED)
    return DPTHybridEncoder(channels_last=False).eval()


def main():
    ours = build_ours()
    with reference_midas() as dpt_depth:
        ref = dpt_depth.DPTDepthModel(path=None, backbone="vitb_rn50_384", non_negative=True).eval()
        ref.load_state_dict(ours.state_dict(), strict=True)
        x, pose = inputs()
        with torch.no_grad():
            p2, p1 = ref(x, pose, 2)
    out = {"seed": SEED, "stride": STRIDE}
    for name, t in (("path_2", p2), ("path_1", p1)):
        out[name] = t.reshape(-1)[::STRIDE].numpy()
        out[name + "_norm"] = float(t.double().norm())
        out[name + "_shape"] = np.array(t.shape)
    np.savez_compressed(os.path.join(HERE, "encoder_golden.npz"), **out)
    print("wrote encoder_golden.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
