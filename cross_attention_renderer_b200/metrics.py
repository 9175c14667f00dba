"""Evaluation metrics of the reference's eval drivers (experiment_scripts/eval_realestate10k.py:74-75,177-199).

* ``psnr``  ``-10 log10(mean((x - y)^2))`` on images in [0, 1] (:74-75,181).
* ``ssim``  restatement of ``skimage.metrics.structural_similarity(x, y, win_size=11, multichannel=True,
  gaussian_weights=True)`` as called at :195, with the behaviour of the pinned scikit-image 0.18.3
  (requirements.txt:13): Gaussian window sigma 1.5 truncated at 3.5 sigma (11 taps), ``scipy.ndimage`` reflect
  borders, population covariances, K1 = 0.01, K2 = 0.03, computed in float64 per channel, mean over the
  image cropped by 5 pixels - and ``data_range = 2`` because float inputs without an explicit range get
  ``dtype_range[float32] = (-1, 1)`` in that version, although the images are in [0, 1].
* ``lpips`` (:190-192) needs the ``lpips`` package and its AlexNet weights; neither is available offline, so
  it is reported as unavailable instead of being approximated.

scikit-image is not installed here, so ``ssim`` is validated against hand-computed cases
(tests/test_metrics.py), not against the package.
"""
import math

import numpy as np
import torch


def psnr(x, y):
    mse = float(torch.mean((x - y) ** 2))
    return -10.0 * math.log10(max(mse, 1e-20))


def _gauss(a, sigma=1.5, truncate=3.5):
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(a, sigma=sigma, truncate=truncate)


def ssim(im1, im2, data_range=2.0, k1=0.01, k2=0.03):
    """im1, im2: (H, W, C) arrays / tensors.  Returns the mean SSIM over channels."""
    a = np.asarray(im1.detach().cpu() if torch.is_tensor(im1) else im1, dtype=np.float64)
    b = np.asarray(im2.detach().cpu() if torch.is_tensor(im2) else im2, dtype=np.float64)
    assert a.shape == b.shape and a.ndim == 3
    win = 2 * int(3.5 * 1.5 + 0.5) + 1                      # 11
    if min(a.shape[:2]) < win:
        raise ValueError("win_size exceeds image extent")
    pad = (win - 1) // 2
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    vals = []
    for ch in range(a.shape[2]):
        x, y = a[..., ch], b[..., ch]
        ux, uy = _gauss(x), _gauss(y)
        uxx, uyy, uxy = _gauss(x * x), _gauss(y * y), _gauss(x * y)
        vx, vy, vxy = uxx - ux * ux, uyy - uy * uy, uxy - ux * uy          # cov_norm = 1 (gaussian weights)
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        vals.append(s[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))


def lpips_available():
    try:
        import lpips  # noqa: F401
        return True
    except Exception:
        return False
