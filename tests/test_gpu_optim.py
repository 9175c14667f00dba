"""FlatAdam / car_adam_step against torch.optim.Adam + clip_grad_norm_ (the reference's optimiser step,
training.py:124-136; Adam(lr, betas=(0.99, 0.999)), train_realestate10k.py:86)."""
import copy

import pytest
import torch

from cross_attention_renderer_b200.optim import FlatAdam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip", [None, 1.0])
def test_flat_adam_matches_torch_adam(clip):
    torch.manual_seed(0)
    dev = "cuda:0"
    net_a = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5)).to(dev)
    net_b = copy.deepcopy(net_a)
    opt_a = torch.optim.Adam(net_a.parameters(), lr=5e-3, betas=(0.99, 0.999))
    opt_b = FlatAdam(net_b.parameters(), lr=5e-3, betas=(0.99, 0.999))
    for step in range(8):
        x = torch.randn(16, 37, device=dev) * (10.0 if step % 2 else 0.1)      # norms above and below the clip
        y = torch.randn(16, 5, device=dev)
        opt_a.zero_grad()
        (net_a(x) - y).pow(2).mean().backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(net_a.parameters(), max_norm=clip)
        opt_a.step()
        opt_b.zero_grad()
        (net_b(x) - y).pow(2).mean().backward()
        opt_b.step(max_grad_norm=clip)
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=1e-7), (step, float((pa - pb).abs().max()))
    sd = opt_b.state_dict()
    ref = opt_a.state_dict()
    assert set(sd) == set(ref) == {"state", "param_groups"}
    for i, st in ref["state"].items():
        # moments: same recurrences, different fp32 op order than torch's foreach kernels
        ea, eq = st["exp_avg"], st["exp_avg_sq"]
        assert torch.allclose(sd["state"][i]["exp_avg"], ea, rtol=1e-3, atol=1e-5 * float(ea.abs().max()))
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], eq, rtol=1e-3, atol=1e-5 * float(eq.abs().max()))
        assert float(sd["state"][i]["step"]) == float(st["step"]) == 8.0
    # the parameters are views of one buffer, the gradients too
    base = opt_b.flat.data_ptr()
    assert all(base <= p.data_ptr() < base + 4 * opt_b.flat.numel() for p in net_b.parameters())
    opt_c = FlatAdam(copy.deepcopy(net_b).parameters(), lr=1e-3)
    opt_c.load_state_dict(sd)
    assert opt_c.step_count == 8 and torch.equal(opt_c.exp_avg, opt_b.exp_avg)
