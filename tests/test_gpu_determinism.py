"""Bit-reproducibility of the CUDA path: repeated renders of the same inputs are identical in every output.
(Round 1 found a TMA bulk-copy staging of Q1 in the tail's phase B that produced run-to-run differences in ~30 of
65 536 rays; it was replaced by per-thread cp.async staging - the suspected cause is a generic-proxy read /
async-proxy write hazard on the aliased shared-memory region without a fence.proxy.async.  This test is the
regression guard for any such cross-proxy race in the fused kernels.)"""
import pytest
import torch

from cross_attention_renderer_b200 import synthetic
from test_gpu_parity import make_model, run_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,P,H", [("fp32", 64, 128), ("bf16", 64, 128), ("fp32", 128, 64)])
def test_repeated_renders_are_bit_identical(precision, P, H):
    b = 1
    inp = synthetic.make_inputs(b, H, H, seed=91, mode="default")
    z = synthetic.make_features(b, H, seed=91)
    sd = synthetic.make_state_dict(seed=91, peaky=True)
    model = make_model(sd, P, H, precision=precision, pixel_val_to_cpu=False)
    ref = run_cuda(model, inp, z)
    keys = ("rgb", "at_wt", "at_wt_max", "depth_ray", "valid_mask", "pixel_val")
    ref = {k: ref[k].clone() for k in keys}
    for rep in range(6):
        out = run_cuda(model, inp, z)
        for k in keys:
            assert torch.equal(out[k], ref[k]), (rep, k, int((out[k] != ref[k]).sum()))
