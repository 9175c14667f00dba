"""Host-side logic that needs no GPU: parameter ABI, packing algebra, module surface."""
import inspect

import pytest
import torch

from cross_attention_renderer_b200 import params, synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer


def test_state_dict_keys_match_reference_abi():
    m = CrossAttentionRenderer(n_view=2, npoints=64)
    sd = m.state_dict()
    spec = params.renderer_param_shapes(2)
    assert set(sd) == set(spec)
    for k, shp in spec.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    # 58 non-encoder tensors, 1 457 955 parameters
    assert len(spec) == 58
    assert sum(v.numel() for v in sd.values()) == 1457955


def test_constructor_signature_matches_reference():
    sig = inspect.signature(CrossAttentionRenderer.__init__)
    names = list(sig.parameters)[1:11]
    assert names == ["no_sample", "no_latent_concat", "no_multiview", "no_high_freq", "model", "uv",
                     "repeat_attention", "n_view", "npoints", "num_hidden_units_phi"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["no_sample"], d["model"], d["repeat_attention"], d["n_view"], d["npoints"],
            d["num_hidden_units_phi"]) == (False, "midas_vit", True, 1, 64, 128)
    fsig = inspect.signature(CrossAttentionRenderer.forward)
    assert list(fsig.parameters)[1:5] == ["input", "z", "val", "debug"]


def test_load_synthetic_state_dict_and_constructor_branches():
    m = CrossAttentionRenderer(n_view=2, npoints=32)
    missing, unexpected = m.load_state_dict(synthetic.make_state_dict(0), strict=False)
    assert not missing and not unexpected
    # every branch of the reference's constructor creates exactly the reference's parameter set
    # (params.py is checked against the unmodified reference in tests/golden/make_golden_nview.py)
    from cross_attention_renderer_b200.params import renderer_param_shapes
    for kw in (dict(n_view=1), dict(n_view=3), dict(n_view=2, no_sample=True), dict(n_view=2, no_latent_concat=True)):
        mm = CrossAttentionRenderer(npoints=16, **kw)
        want = renderer_param_shapes(kw["n_view"], no_latent_concat=kw.get("no_latent_concat", False))
        got = {k: tuple(v.shape) for k, v in mm.state_dict().items() if not k.startswith("encoder.")}
        assert got == {k: tuple(v) for k, v in want.items()}, kw
        assert mm.general
    assert not m.general
    with pytest.raises(NotImplementedError):
        CrossAttentionRenderer(n_view=4)
    with pytest.raises(NotImplementedError):
        CrossAttentionRenderer(n_view=3, no_sample=True)


def test_packing_algebra():
    from cross_attention_renderer_b200.packing import PackedWeights
    sd = synthetic.make_state_dict(5)
    pw = PackedWeights(sd)
    e1 = pw.m["enc1"]
    assert (e1.N, e1.K) == (576, 592)
    assert torch.equal(e1.f32[:, :579], sd["query_encode_latent.weight"].reshape(576, 579))
    assert float(e1.f32[:, 579:].abs().max()) == 0.0
    # hi + lo reproduces fp32 to ~2^-17 relative
    rec = e1.hi.float() + e1.lo.float()
    assert float((rec - e1.f32).abs().max() / e1.f32.abs().max()) < 2 ** -15
    # query_repeat_embed split: W·[g|local] == Wg·g + Wl·local
    w = sd["query_repeat_embed.weight"].reshape(128, 144)
    g, loc = torch.randn(7, 128), torch.randn(7, 16)
    full = torch.cat([g, loc], -1) @ w.T
    split = g @ pw.m["rep1_g"].f32.T + loc @ pw.m["rep1_loc"].f32.T
    assert torch.allclose(full, split, atol=1e-5)
    # lin_z fold: W·[z|z] == (Wa+Wb)·z
    wz = sd["phi.lin_z.1.weight"]
    zz = torch.randn(5, 288)
    assert torch.allclose(torch.cat([zz, zz], -1) @ wz.T, zz @ pw.m["phi_z1"].f32.T, atol=1e-5)


def test_training_pack_skips_the_inference_folds_and_precision_selection():
    """The training path packs the plain layers only (the composed matrices feed the fused inference kernels),
    and picks its GEMM arithmetic from the module's precision (backward separately on request)."""
    from cross_attention_renderer_b200 import _lib
    from cross_attention_renderer_b200.packing import PackedGrads, PackedWeights
    sd = synthetic.make_state_dict(5)
    full, plain = PackedWeights(sd), PackedWeights(sd, folds=False)
    assert set(full.m) - set(plain.m) == {"kv_fold", "kv_fold64", "rowb_fold", "phi_pack"}
    for k in plain.m:
        assert torch.equal(plain.m[k].f32, full.m[k].f32)
    w = plain.c_struct()
    assert not w.kv_fold.hi and not w.phi_pack.hi and w.enc1.hi and (w.enc1.N, w.enc1.K) == (576, 592)
    assert set(PackedGrads(plain).g) == set(PackedGrads(full).g)              # gradients exist for the plain layers only
    m = CrossAttentionRenderer(n_view=2, npoints=8, precision="fp32")
    assert m._train_precision() == _lib.PREC_FP32_3XBF16 == m._train_precision(backward=True)
    m.backward_precision = "fp32_simt"
    assert m._train_precision() == _lib.PREC_FP32_3XBF16 and m._train_precision(backward=True) == _lib.PREC_FP32_SIMT
    assert CrossAttentionRenderer(n_view=2, npoints=8, precision="fp32_simt")._train_precision() == _lib.PREC_FP32_SIMT
    assert CrossAttentionRenderer(n_view=2, npoints=8, precision="bf16")._train_precision() == _lib.PREC_FP32_3XBF16


def test_synthetic_inputs_are_deterministic():
    a = synthetic.make_inputs(2, 32, 8, seed=3, mode="mixed")
    b = synthetic.make_inputs(2, 32, 8, seed=3, mode="mixed")
    assert torch.equal(a["context"]["cam2world"], b["context"]["cam2world"])
    assert a["query"]["uv"].shape == (2, 1, 64, 2)
    z = synthetic.make_features(2, 32, seed=3)
    assert [tuple(t.shape) for t in z] == [(4, 256, 8, 8), (4, 256, 16, 16), (4, 64, 32, 32)]
