"""A small stand-in for the reference's image encoder (NOT the reference's DPT-hybrid ViT).

The reference encoder (``midas/dpt_depth.py:67-89`` on timm 0.5.4) is outside the hot path
 and its dependency is not available offline.  The drivers under
``experiment_scripts/`` still need *an* encoder that honours the same contract so that
``get_z`` -> ``forward`` -> loss -> backward can be exercised end to end, including the
gradient that the renderer's backward pass hands to the feature maps:

    forward(rgb (b*n,3,H,W) normalised, cam2world_encode (b*n,16), n_view)
        -> [path_2 (b*n,256,H/4,W/4), path_1 (b*n,256,H/2,W/2)]

Plain torch convolutions (cuDNN): plumbing, not the product.  Pass the reference's real
``encoder`` module to ``CrossAttentionRenderer(encoder=...)`` to reproduce its results.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class StandInEncoder(nn.Module):
    def __init__(self, width=64):
        super().__init__()
        self.stem = nn.Conv2d(3, width, 7, stride=2, padding=3)          # H/2
        self.mid = nn.Conv2d(width, 2 * width, 3, stride=2, padding=1)   # H/4
        self.pose_embed = nn.Linear(16, 2 * width)                       # role of vit_models.py:80
        self.mix = nn.Conv2d(4 * width, 2 * width, 1)                    # cross-view exchange at H/4
        self.out4 = nn.Conv2d(2 * width, 256, 3, padding=1)
        self.up = nn.Conv2d(2 * width + width, 256, 3, padding=1)

    def forward(self, rgb, cam2world_encode, n_view):
        bn = rgb.shape[0]
        x2 = F.relu(self.stem(rgb))
        x4 = F.relu(self.mid(x2)) + self.pose_embed(cam2world_encode)[:, :, None, None]
        # every view sees the mean of the views of its scene (stand-in for the joint-token attention
        # of midas/vit.py:185-186)
        b = bn // n_view
        scene = x4.reshape(b, n_view, *x4.shape[1:]).mean(dim=1, keepdim=True).expand(-1, n_view, -1, -1, -1)
        x4 = F.relu(self.mix(torch.cat([x4, scene.reshape_as(x4)], dim=1)))
        path_2 = self.out4(x4)
        up = F.interpolate(x4, size=x2.shape[-2:], mode="bilinear", align_corners=False)
        path_1 = self.up(torch.cat([up, x2], dim=1))
        return [path_2, path_1]
