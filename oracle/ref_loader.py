"""ORACLE — test infrastructure, NOT product code.

Imports and runs the UNMODIFIED reference from ``oracle/_ref`` (built by ``oracle/build_ref.py``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference legs may import this.

The reference imports timm / matplotlib / midas at module scope (models.py:6,14; utils/util.py:1);
none of them is on the arithmetic path of ``CrossAttentionRenderer.forward(input, z=z)``, so they are
replaced by empty stub modules (same stubs as ``tests/golden/make_golden.py``).  On a CUDA device the
reference's hard-coded ``.cuda()`` calls (geometry.py:320,398) are real; for a host run they are made
no-ops for the duration of the call (``host_mode``)."""
import contextlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mod = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "models.py"))


def load():
    """The reference's ``models`` module (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.colors", "timm", "timm.models", "timm.models.layers",
                 "timm.models.layers.std_conv", "midas", "midas.dpt_depth", "midas.midas_net",
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
        pass

    class _DPT(torch.nn.Module):           # encoder placeholder: forward(z=...) never calls it
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
    # the reference's top-level module names (models, geometry, epipolar, utils, ...) resolve to oracle/_ref
    for name in ("models", "geometry", "epipolar", "encoder", "resnet_block_fc", "utils", "utils.util", "utils.pixel_util"):
        sys.modules.pop(name, None)
    sys.path.insert(0, REF_DIR)
    try:
        import models as ref_models        # noqa: E402
    finally:
        sys.path.remove(REF_DIR)
    _mod = ref_models
    return _mod


@contextlib.contextmanager
def host_mode():
    """Run the reference on the host: ``Tensor.cuda`` is a no-op inside the block."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def strict_fp32():
    """TF32 off for convolutions and matmuls (cuDNN convs default to TF32): the reference as an fp32 oracle."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m


def build_model(sd, H, P, n_view=2, device="cpu", **kw):
    """Reference ``CrossAttentionRenderer`` with the given renderer weights (encoder.* keys stay missing)."""
    ref_models = load()
    torch.manual_seed(0)
    m = ref_models.CrossAttentionRenderer(model="midas_vit", n_view=n_view, npoints=P, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("encoder.") for k in missing), missing
    m.eval()
    m.H = m.W = H
    return m.to(device)


def render(m, inp, z, chunk_rays=None):
    """``m(inp, z=z)`` over ray chunks like the reference's own drivers do (eval_realestate10k.py:144-159,
    render_realestate10k_traj.py:118-130): returns the concatenated out_dict entries."""
    uv = inp["query"]["uv"]
    R = uv.shape[2]
    chunk = chunk_rays or R
    parts = []
    with torch.no_grad():
        for r0 in range(0, R, chunk):
            q = dict(inp["query"])
            q["uv"] = uv[:, :, r0:r0 + chunk]
            parts.append(m({"context": inp["context"], "query": q}, z=z))
    if len(parts) == 1:
        return parts[0]
    out = {}
    ray_dim = {"rgb": 2, "valid_mask": 1, "depth_ray": 1, "at_wt": 1, "at_wt_max": 1, "pixel_val": 1, "coords": 1}
    for k, d in ray_dim.items():
        out[k] = torch.cat([p[k] for p in parts], dim=d)
    return out
