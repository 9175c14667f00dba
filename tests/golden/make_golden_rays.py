#!/usr/bin/env python
"""Golden vectors for the epipolar clipping on DEGENERATE rays, produced by calling the unmodified
reference's ``epipolar.project_rays`` directly (CPU): origin at the camera centre, behind the camera, on
the z = 0 plane; directions parallel to the image plane, through the principal point, along an image edge,
with zero components (the +-inf / NaN paths of epipolar.py:99-116 and the all-invalid tie-break of :142).

    python tests/golden/make_golden_rays.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference                            # noqa: E402


def crafted():
    H = 64
    f, c = 225.0 * H / 256.0, H / 2.0
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = f
    K[0, 2] = K[1, 2] = c
    g = torch.Generator().manual_seed(123)
    origins = torch.tensor([
        [0.0, 0.0, 0.0],          # at the camera centre (epipolar.py:213-215)
        [0.0, 0.0, -1.0],         # behind the camera
        [0.3, -0.2, 0.0],         # on the z = 0 plane
        [0.5, 0.1, 2.0],          # in front, inside the frustum
        [5.0, 0.0, 1.0],          # in front, far outside the frustum
        [1e-7, 0.0, 0.0],         # almost at the centre (norm < 1e-6)
        [0.0, 0.0, 1e-7],         # z just below the 1e-6 threshold
        [-0.4, 0.3, -0.5],
    ])
    dirs = [
        [0.0, 0.0, 1.0], [0.0, 0.0, -1.0],                # along the optical axis
        [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.7071067690849304, 0.7071067690849304, 0.0],   # parallel to the image plane
        [0.0, 0.6, 0.8], [0.6, 0.0, 0.8],                  # one zero component
        [(0.0 - c) / f, (0.0 - c) / f, 1.0], [(H - c) / f, 0.0, 1.0],   # towards a corner / an edge midpoint
        [0.1, -0.05, 1.0], [-0.3, 0.2, -1.0],
    ]
    d = torch.tensor(dirs)
    d = d / d.norm(dim=-1, keepdim=True)
    d = torch.cat([d, torch.nn.functional.normalize(torch.randn(20, 3, generator=g), dim=-1)], 0)
    return H, K, origins, d


def main():
    import_reference()
    from epipolar import project_rays
    H, K, origins, d = crafted()
    B, R = origins.shape[0], d.shape[0]
    Kn = K.clone()
    Kn[:2, :] = Kn[:2, :] / H                                       # models.py:228
    out = project_rays(origins[:, None, :].expand(B, R, 3).contiguous(), d[None].expand(B, R, 3).contiguous(),
                       torch.eye(4)[None].expand(B, 4, 4).contiguous(), Kn[None].expand(B, 4, 4).contiguous())
    rec = {"H": np.array(H), "K": K.numpy(), "origins": origins.numpy(), "dirs": d.numpy(),
           "xy_min": out["xy_min"].numpy(), "xy_max": out["xy_max"].numpy(),
           "overlaps": out["overlaps_image"].numpy()}
    path = os.path.join(HERE, "rays_degenerate.npz")
    np.savez_compressed(path, **rec)
    print({k: v.shape for k, v in rec.items()}, "%.1f KB" % (os.path.getsize(path) / 1024),
          "overlapping:", int(out["overlaps_image"].sum()), "of", B * R,
          "non-finite xy:", int((~torch.isfinite(out["xy_min"])).sum() + (~torch.isfinite(out["xy_max"])).sum()))


if __name__ == "__main__":
    main()
