"""The image encoder (cross_attention_renderer_b200/encoder.py).

* against golden vectors produced by the UNMODIFIED reference DPT wrapper (tests/golden/make_golden_encoder.py) and,
  where /root/reference exists, against that wrapper live: outputs and ``state_dict`` keys;
* the timm 0.5.4 building blocks (absent here) against independent statements of their published definitions;
* ``get_z`` through the drop-in module, NHWC take-over of channels_last maps.
CPU only: the encoder is plain torch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cross_attention_renderer_b200 import encoder as enc
from cross_attention_renderer_b200 import packing, synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_encoder as gold                                    # noqa: E402


@pytest.fixture(scope="module")
def ours():
    return gold.build_ours()


def test_encoder_matches_reference_golden(ours):
    d = np.load(os.path.join(HERE, "golden", "encoder_golden.npz"))
    assert int(d["seed"]) == gold.SEED
    x, pose = gold.inputs()
    with torch.no_grad():
        p2, p1 = ours(x, pose, 2)
    for name, t in (("path_2", p2), ("path_1", p1)):
        assert tuple(t.shape) == tuple(d[name + "_shape"])
        got = t.reshape(-1)[::int(d["stride"])]
        ref = torch.from_numpy(d[name])
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), name
        assert abs(float(t.double().norm()) - float(d[name + "_norm"])) <= 1e-5 * float(d[name + "_norm"])
    # channels_last execution gives the same maps, in NHWC memory order
    m2 = gold.build_ours()
    m2.channels_last = True
    with torch.no_grad():
        q2, q1 = m2(x, pose, 2)
    assert q1.is_contiguous(memory_format=torch.channels_last) and packing.nhwc_view(q1) is not None
    assert float((q1 - p1).abs().max()) <= 2e-5 * float(p1.abs().max())
    assert float((q2 - p2).abs().max()) <= 2e-5 * float(p2.abs().max())


@pytest.mark.skipif(not os.path.isdir(gold.REF), reason="reference sources not present (GPU box)")
def test_encoder_matches_reference_wrapper_live(ours):
    """The reference's DPTDepthModel (its read-outs, reassemble stacks, scratch convolutions, fusion blocks, hooks,
    forward_vit and multi-view forward_flex) built around this file's ViT: identical state_dict keys both ways and
    identical outputs on every element."""
    with gold.reference_midas() as dpt_depth:
        ref = dpt_depth.DPTDepthModel(path=None, backbone="vitb_rn50_384", non_negative=True).eval()
        assert set(ref.state_dict()) == set(ours.state_dict())
        for k, v in ref.state_dict().items():
            assert v.shape == ours.state_dict()[k].shape, k
        ref.load_state_dict(ours.state_dict(), strict=True)
        x, pose = gold.inputs()
        with torch.no_grad():
            r2, r1 = ref(x, pose, 2)
            p2, p1 = ours(x, pose, 2)
    assert float((p2 - r2).abs().max()) <= 2e-5 * float(r2.abs().max())
    assert float((p1 - r1).abs().max()) <= 2e-5 * float(r1.abs().max())


def test_views_attend_jointly_and_pose_matters(ours):
    """The multi-view part of forward_flex (midas/vit.py:176-185): tokens of the views of one scene share the
    attention, so a view's maps change when its partner changes; the pose embedding reaches every token."""
    x, pose = gold.inputs()
    with torch.no_grad():
        a = ours(x, pose, 2)[1]
        x2 = x.clone()
        x2[1] = torch.randn_like(x2[1])
        b = ours(x2, pose, 2)[1]
        c = ours(x, pose * 0, 2)[1]
        solo = ours(x, pose, 1)[1]                 # each image its own scene: no cross-view attention
        solo2 = ours(x2, pose, 1)[1]
    assert float((a[0] - b[0]).abs().max()) > 1e-4            # view 0 sees the change of view 1
    assert float((solo[0] - solo2[0]).abs().max()) == 0.0      # ... but not when the views are separate scenes
    assert float((a - c).abs().max()) > 1e-4


def test_timm_blocks_against_their_definitions():
    torch.manual_seed(0)
    # StdConv2dSame: per-filter weight standardisation (biased variance) + TensorFlow SAME padding
    for k, s, n in ((7, 2, 37), (3, 2, 16), (3, 1, 9), (1, 2, 8), (1, 1, 5)):
        conv = enc.StdConv2dSame(5, 6, k, stride=s, eps=1e-8)
        x = torch.randn(2, 5, n, n + 3)
        w = conv.weight
        mu = w.mean(dim=(1, 2, 3), keepdim=True)
        var = w.var(dim=(1, 2, 3), keepdim=True, unbiased=False)
        wn = (w - mu) / torch.sqrt(var + 1e-8)
        pads = []
        for size in (x.shape[-1], x.shape[-2]):                 # F.pad order: last dim first
            out = -(-size // s)
            tot = max((out - 1) * s + k - size, 0)
            pads += [tot // 2, tot - tot // 2]
        ref = F.conv2d(F.pad(x, pads), wn, None, s)
        got = conv(x)
        assert got.shape == ref.shape and got.shape[-1] == -(-x.shape[-1] // s)
        assert torch.allclose(got, ref, atol=1e-5), (k, s)
    # GroupNormAct = GroupNorm(32) (+ ReLU)
    gn = enc.GroupNormAct(64)
    x = torch.randn(2, 64, 5, 5)
    assert torch.allclose(gn(x), F.relu(F.group_norm(x, 32, gn.weight, gn.bias, 1e-5)))
    assert float(enc.GroupNormAct(64, apply_act=False)(x).min()) < 0
    # MaxPool2dSame: -inf padding, output = ceil(n / 2)
    x = torch.randn(1, 3, 7, 10)
    mp = enc.MaxPool2dSame(3, 2)(x)
    assert mp.shape[-2:] == (4, 5)
    assert torch.equal(mp, F.max_pool2d(F.pad(x, [0, 1, 1, 1], value=-float("inf")), 3, 2))
    # attention block: softmax(q k^T / sqrt(d)) v with fused qkv, pre-norm residuals, GELU MLP
    blk = enc._Block(48, 4)
    t = torch.randn(2, 11, 48)
    y = blk.norm1(t)
    qkv = blk.attn.qkv(y).reshape(2, 11, 3, 4, 12).permute(2, 0, 3, 1, 4)
    att = ((qkv[0] @ qkv[1].transpose(-2, -1)) * 12 ** -0.5).softmax(-1)
    a = blk.attn.proj((att @ qkv[2]).transpose(1, 2).reshape(2, 11, 48))
    r = t + a
    r = r + blk.mlp.fc2(F.gelu(blk.mlp.fc1(blk.norm2(r))))
    assert torch.allclose(blk(t), r, atol=1e-5)
    # the (3, 4, 9) backbone: strides 4 / 8 / 16, widths 256 / 512 / 1024, the key names timm gives them
    bb = enc.ResNetV2Backbone()
    keys = set(bb.state_dict())
    for k in ("stem.conv.weight", "stem.norm.weight", "stages.0.blocks.0.downsample.conv.weight",
              "stages.0.blocks.0.downsample.norm.bias", "stages.2.blocks.8.conv3.weight", "stages.1.blocks.3.norm2.weight"):
        assert k in keys, k
    assert not any(k.startswith("norm.") or k.startswith("head.") for k in keys)
    with torch.no_grad():
        s = bb.stem(torch.randn(1, 3, 64, 64))
        s0 = bb.stages[0](s)
        s1 = bb.stages[1](s0)
        s2 = bb.stages[2](s1)
    assert s0.shape == (1, 256, 16, 16) and s1.shape == (1, 512, 8, 8) and s2.shape == (1, 1024, 4, 4)


def test_get_z_through_the_module_and_nhwc_takeover():
    torch.manual_seed(3)
    m = CrossAttentionRenderer(n_view=2, npoints=8, encoder="dpt_hybrid").eval()
    assert sum(p.numel() for p in m.encoder.parameters()) == 123603177
    assert any(k.startswith("encoder.pretrained.model.blocks.11.mlp.fc2") for k in m.state_dict())
    inp = synthetic.make_inputs(1, 256, 4, seed=2)
    inp["context"]["rgb"] = torch.rand(1, 2, 256, 256, 3) * 2 - 1
    with torch.no_grad():
        z = m.get_z(inp)
    assert [tuple(t.shape) for t in z] == [(2, 256, 64, 64), (2, 256, 128, 128), (2, 64, 256, 256)]
    for t in z:                                                # NHWC in memory: the renderer's packed layout as is
        v = packing.nhwc_view(t)
        assert v is not None and v.data_ptr() == t.data_ptr() and v.shape[-1] == t.shape[1]
    assert packing.nhwc_view(torch.randn(2, 8, 4, 4)) is None  # NCHW maps still go through car_pack_features
    m.no_high_freq = True
    with torch.no_grad():
        assert float(m.get_z(inp)[2].abs().max()) == 0.0       # models.py:183-184
    with pytest.raises(ValueError):
        CrossAttentionRenderer(n_view=2, encoder="resnet")


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_channels_last_maps_render_identically(precision):
    """Maps handed over in NHWC memory order (the encoder's output) are taken as the packed layout without the
    car_pack_features pass: same bits as rendering from the NCHW copies (bf16: the same round-to-nearest conversion)."""
    dev = "cuda:0"
    b, H, Ht, P = 1, 64, 16, 64
    inp = synthetic.to_device(synthetic.make_inputs(b, H, Ht, seed=41), dev)
    z = [t.to(dev) for t in synthetic.make_features(b, H, seed=41)]
    zc = [t.contiguous(memory_format=torch.channels_last) for t in z]
    assert all(packing.nhwc_view(t) is not None for t in zc) and all(packing.nhwc_view(t) is None for t in z)
    m = CrossAttentionRenderer(n_view=2, npoints=P, precision=precision).to(dev).eval()
    m.load_state_dict(synthetic.make_state_dict(seed=41), strict=False)
    m.H = m.W = H
    with torch.no_grad():
        a = m(inp, z=z)
        m.release_features()
        c = m(inp, z=zc)
    for k in ("rgb", "depth_ray", "at_wt", "valid_mask"):
        assert torch.equal(a[k], c[k]), k


@pytest.mark.skipif(not os.path.isdir(gold.REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("nviews,flags", [(2, {}), (3, {"no_multiview": True, "no_high_freq": True})])
def test_get_z_matches_the_reference_get_z(nviews, flags):
    """``CrossAttentionRenderer.get_z`` of the UNMODIFIED reference (models.py:148-188: relative poses, ImageNet
    normalisation, encoder call, conv_map, no_multiview / no_high_freq) with the reference's own DPT wrapper as its
    encoder, against this module's get_z (2 and 3 context views; n_view = 1 passes as well, left out for run time)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    torch.manual_seed(7)
    ours_m = CrossAttentionRenderer(n_view=nviews, npoints=8, encoder="dpt_hybrid", **flags).eval()
    sd = synthetic.make_state_dict(seed=7, n_view=nviews)
    ref_m = ref_loader.build_model(sd, 256, 8, n_view=nviews, **flags)
    with gold.reference_midas() as dpt_depth:
        ref_enc = dpt_depth.DPTDepthModel(path=None, backbone="vitb_rn50_384", non_negative=True).eval()
        ref_enc.load_state_dict(ours_m.encoder.state_dict(), strict=True)
        ref_m.encoder = ref_enc
        ref_m.conv_map.load_state_dict(ours_m.conv_map.state_dict())
        inp = synthetic.make_inputs(1, 256, 4, seed=5, n_ctx=nviews)
        inp["context"]["rgb"] = torch.rand(1, nviews, 256, 256, 3, generator=torch.Generator().manual_seed(9)) * 2 - 1
        with torch.no_grad():
            z_ref = ref_m.get_z(inp)
            z = ours_m.get_z(inp)
    assert len(z) == len(z_ref) == 3 and (ref_m.H, ref_m.W) == (ours_m.H, ours_m.W) == (256, 256)
    for a, r in zip(z, z_ref):
        assert a.shape == r.shape
        scale = float(r.abs().max())
        assert float((a - r).abs().max()) <= 2e-5 * max(scale, 1e-6)
    if flags.get("no_high_freq"):
        assert float(z[2].abs().max()) == 0.0
