"""Eval metrics and tensorboard summaries (host logic; reference eval_realestate10k.py:177-199, summaries.py)."""
import numpy as np
import pytest
import torch

from cross_attention_renderer_b200 import metrics, summaries


def test_ssim_properties_and_hand_case():
    g = np.random.default_rng(0)
    a = g.random((40, 48, 3))
    assert abs(metrics.ssim(a, a) - 1.0) < 1e-12
    b = np.clip(a + 0.1 * g.standard_normal(a.shape), 0, 1)
    s_ab, s_ba = metrics.ssim(a, b), metrics.ssim(b, a)
    assert abs(s_ab - s_ba) < 1e-12 and 0.0 < s_ab < 1.0
    assert metrics.ssim(a, np.clip(a + 0.3 * g.standard_normal(a.shape), 0, 1)) < s_ab      # more noise, lower score
    # constant images: zero variances -> S = (2 ux uy + C1) / (ux^2 + uy^2 + C1) with C1 = (0.01 * 2)^2
    x, y = np.full((32, 32, 1), 0.25), np.full((32, 32, 1), 0.75)
    c1 = (0.01 * 2.0) ** 2
    assert abs(metrics.ssim(x, y) - (2 * 0.25 * 0.75 + c1) / (0.25 ** 2 + 0.75 ** 2 + c1)) < 1e-12
    with pytest.raises(ValueError):
        metrics.ssim(np.zeros((8, 8, 3)), np.zeros((8, 8, 3)))
    t = torch.rand(16, 16, 3)
    assert abs(metrics.psnr(t, t + 0.1) - 20.0) < 1e-4


def test_jet_colormap_anchor_points():
    c = summaries.jet(np.array([0.0, 0.5, 1.0]))
    assert np.allclose(c[0], [0.0, 0.0, 0.5], atol=1e-6)                 # dark blue
    assert np.allclose(c[2], [0.5, 0.0, 0.0], atol=1e-6)                 # dark red
    assert c[1][1] > 0.99 and 0.4 < c[1][0] < 0.6 and 0.4 < c[1][2] < 0.6     # green-ish middle
    assert summaries.jet(np.zeros((2, 3, 4))).shape == (2, 3, 4, 3)


class _Writer:
    def __init__(self):
        self.scalars, self.images = {}, {}

    def add_scalar(self, tag, v, it):
        self.scalars[tag] = float(v)

    def add_image(self, tag, img, it):
        self.images[tag] = np.asarray(img)


def test_img_summaries_tags_and_shapes():
    b, n, H, P = 2, 2, 16, 8
    R = H * H
    g = torch.Generator().manual_seed(0)
    out = {"rgb": torch.rand(b, 1, R, 3, generator=g) * 2 - 1, "depth_ray": torch.rand(b, R, 1, generator=g) * 10,
           "at_wt": torch.softmax(torch.randn(b * n, R, P, generator=g), -1),
           "at_wt_max": torch.randint(0, P, (b * n, R, 1), generator=g),
           "pixel_val": torch.rand(b * n, R, P, 2, generator=g) * 2 - 1,
           "uv": torch.stack(torch.meshgrid(torch.arange(H), torch.arange(H), indexing="xy"), -1).reshape(1, 1, R, 2).expand(b, 1, R, 2).float()}
    inp = {"context": {"rgb": torch.rand(b, n, H, H, 3, generator=g) * 2 - 1},
           "query": {"rgb": torch.rand(b, 1, R, 3, generator=g) * 2 - 1}}
    w = _Writer()
    summaries.img_summaries(None, inp, None, None, out, w, 3, prefix="val_", img_shape=(H, H), n_view=n)
    assert set(w.images) == {"val_predictions", "val_depth_images", "val_context_images", "val_query_images", "val_epipolar_line"}
    assert set(w.scalars) == {"val_ent", "val_out_min", "val_out_max", "val_trgt_min", "val_trgt_max"}
    assert w.images["val_predictions"].shape[0] == 3 and w.images["val_epipolar_line"].shape[0] == 3
    assert 0.0 < w.scalars["val_ent"] < np.log(P) + 1e-3
