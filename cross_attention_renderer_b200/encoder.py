"""Image encoder of the reference: the multi-view DPT-hybrid that ``get_z`` runs once per
scene (reference ``models.py:81-96,148-188``, ``midas/dpt_depth.py:26-91``, ``midas/vit.py:57-200,389-541``,
``midas/blocks.py:51-342``, ``vit_models.py:8-204``), without timm.

Plumbing, not a hot-path kernel: plain torch modules (cuDNN / cuBLAS underneath), run once per scene while the
renderer runs once per ray.  What is restated here and how it is pinned:

* the reference's own code - the multi-view ``forward_flex`` (tokens of the n views of a scene attend jointly,
  pose embedding added to every token), the read-out / reassemble blocks, the scratch convolutions and the four
  fusion blocks - is pinned against the UNMODIFIED reference modules (``tests/test_encoder.py`` builds the
  reference's ``DPTDepthModel`` around this file's ViT and compares outputs and ``state_dict`` keys);
* the parts the reference takes from **timm 0.5.4** (absent here; ``requirements.txt:16``) - ResNetV2 (3, 4, 9)
  stem + stages with weight-standardised "same"-padded convolutions and GroupNorm(32), ``HybridEmbed``, the ViT
  ``Block`` - are restated from timm's published code; parity of these is anchored on their definitions only
  (PARITY UNPINNED for the timm internals: timm cannot be imported in this environment).

Module and parameter names follow the reference's ``state_dict`` (``encoder.pretrained.model.*``,
``encoder.pretrained.act_postprocess{3,4}.*``, ``encoder.scratch.*``) so its checkpoints load."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------
# timm 0.5.4 building blocks (layers/padding.py, layers/std_conv.py, layers/norm_act.py, layers/pool2d_same.py)
# ---------------------------------------------------------------------------------------------------------
def _same_pad_amount(x, k, s, d):
    return max((math.ceil(x / s) - 1) * s + (k - 1) * d + 1 - x, 0)


def _pad_same(x, k, s, d=(1, 1), value=0.0):
    """TensorFlow 'SAME' padding computed from the input size (asymmetric: the extra pixel goes right / bottom)."""
    ih, iw = x.shape[-2:]
    ph, pw = _same_pad_amount(ih, k[0], s[0], d[0]), _same_pad_amount(iw, k[1], s[1], d[1])
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)
    return x


class StdConv2dSame(nn.Conv2d):
    """Convolution with weight standardisation and 'SAME' padding: static symmetric padding when stride == 1,
    input-dependent padding otherwise (timm ``get_padding_value``)."""

    def __init__(self, cin, cout, kernel_size, stride=1, dilation=1, groups=1, bias=False, eps=1e-6):
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        s = stride if isinstance(stride, int) else stride[0]
        static = s == 1 and (dilation * (k - 1)) % 2 == 0
        pad = ((s - 1) + dilation * (k - 1)) // 2 if static else 0
        super().__init__(cin, cout, kernel_size, stride=stride, padding=pad, dilation=dilation, groups=groups, bias=bias)
        self.same_pad = not static
        self.eps = eps

    def forward(self, x):
        if self.same_pad:
            x = _pad_same(x, self.kernel_size, self.stride, self.dilation)
        w = F.batch_norm(self.weight.reshape(1, self.out_channels, -1), None, None, training=True, momentum=0.0,
                         eps=self.eps).reshape_as(self.weight)         # (w - mean) / sqrt(biased var + eps) per filter
        return F.conv2d(x, w, self.bias, self.stride, self.padding, self.dilation, self.groups)


class GroupNormAct(nn.GroupNorm):
    def __init__(self, num_channels, num_groups=32, eps=1e-5, apply_act=True):
        super().__init__(num_groups, num_channels, eps=eps, affine=True)
        self.apply_act = apply_act

    def forward(self, x):
        x = F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return F.relu(x) if self.apply_act else x


class MaxPool2dSame(nn.Module):
    def __init__(self, kernel_size=3, stride=2):
        super().__init__()
        self.k, self.s = (kernel_size, kernel_size), (stride, stride)

    def forward(self, x):
        return F.max_pool2d(_pad_same(x, self.k, self.s, value=-float("inf")), self.k, self.s)


# ---------------------------------------------------------------------------------------------------------
# timm 0.5.4 ResNetV2 (models/resnetv2.py), the non-pre-activation variant `_resnetv2((3, 4, 9))` builds
# (models/vision_transformer_hybrid.py): stem conv 7x7/2 + GroupNorm + max-pool, three stages of bottlenecks
# ---------------------------------------------------------------------------------------------------------
def _conv(cin, cout, k, stride=1):
    return StdConv2dSame(cin, cout, k, stride=stride, eps=1e-8)


class _DownsampleConv(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = _conv(cin, cout, 1, stride)
        self.norm = GroupNormAct(cout, apply_act=False)

    def forward(self, x):
        return self.norm(self.conv(x))


class _Bottleneck(nn.Module):
    def __init__(self, cin, cout, stride, downsample):
        super().__init__()
        mid = cout // 4
        self.downsample = _DownsampleConv(cin, cout, stride) if downsample else None
        self.conv1 = _conv(cin, mid, 1)
        self.norm1 = GroupNormAct(mid)
        self.conv2 = _conv(mid, mid, 3, stride)
        self.norm2 = GroupNormAct(mid)
        self.conv3 = _conv(mid, cout, 1)
        self.norm3 = GroupNormAct(cout, apply_act=False)

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        x = self.norm1(self.conv1(x))
        x = self.norm2(self.conv2(x))
        x = self.norm3(self.conv3(x))
        return F.relu(x + shortcut)


class _Stage(nn.Module):
    def __init__(self, cin, cout, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(*[_Bottleneck(cin if i == 0 else cout, cout, stride if i == 0 else 1, i == 0)
                                      for i in range(depth)])

    def forward(self, x):
        return self.blocks(x)


class ResNetV2Backbone(nn.Module):
    def __init__(self, layers=(3, 4, 9), channels=(256, 512, 1024), in_chans=3, stem_chs=64):
        super().__init__()
        self.stem = nn.Sequential()
        self.stem.add_module("conv", _conv(in_chans, stem_chs, 7, 2))
        self.stem.add_module("norm", GroupNormAct(stem_chs))
        self.stem.add_module("pool", MaxPool2dSame(3, 2))
        self.stages = nn.Sequential()
        prev = stem_chs
        for i, (d, c) in enumerate(zip(layers, channels)):
            self.stages.add_module(str(i), _Stage(prev, c, 1 if i == 0 else 2, d))
            prev = c
        self.num_features = prev
        self.norm = nn.Identity()                      # only the pre-activation variant normalises here

    def forward(self, x):
        return self.norm(self.stages(self.stem(x)))


class _HybridEmbed(nn.Module):
    """timm ``HybridEmbed`` with patch_size 1: the CNN feature map (stride 16) projected to the token width."""

    def __init__(self, backbone, img_size=384, embed_dim=768):
        super().__init__()
        self.backbone = backbone
        self.proj = nn.Conv2d(backbone.num_features, embed_dim, kernel_size=1, stride=1)
        self.num_patches = (img_size // 16) ** 2

    def forward(self, x):
        return self.proj(self.backbone(x)).flatten(2).transpose(1, 2)


# ---------------------------------------------------------------------------------------------------------
# timm 0.5.4 ViT block (models/vision_transformer.py: Attention, Mlp, Block)
# ---------------------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4).unbind(0)
        x = F.scaled_dot_product_attention(q, k, v, scale=self.scale)       # softmax(q k^T * scale) v
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


# ---------------------------------------------------------------------------------------------------------
# The reference's multi-view ViT (vit_models.py:8-90) with its forward_flex (midas/vit.py:124-200)
# ---------------------------------------------------------------------------------------------------------
class VisionTransformerMultiView(nn.Module):
    """R50 + ViT-B/16 hybrid with the reference's additions: ``pose_embed`` (Linear(16, 768) on the flattened
    relative camera-to-world matrix, added to every token of the view) and joint attention over the tokens of all
    views of a scene.  ``pos_embed_second`` and ``head`` exist for ``state_dict`` parity only (the reference
    resizes the former and never adds it, midas/vit.py:130-132,176)."""

    def __init__(self, img_size=384, embed_dim=768, depth=12, num_heads=12, num_classes=1000, stem_eps=1e-6):
        super().__init__()
        backbone = ResNetV2Backbone()
        # models.py:94: the constructor swaps the stem convolution for a fresh StdConv2dSame (default eps 1e-6)
        backbone.stem.conv = StdConv2dSame(3, 64, 7, stride=2, bias=False, eps=stem_eps)
        self.patch_embed = _HybridEmbed(backbone, img_size, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_embed_second = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.pre_logits = nn.Identity()
        self.pose_embed = nn.Linear(16, embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.start_index = 1
        self.patch_size = [16, 16]
        for p in (self.pos_embed, self.pos_embed_second, self.cls_token):
            nn.init.trunc_normal_(p, std=0.02)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def _resize_pos_embed(self, posemb, gh, gw):                          # midas/vit.py:105-121
        tok, grid = posemb[:, :self.start_index], posemb[0, self.start_index:]
        g = int(math.sqrt(len(grid)))
        grid = grid.reshape(1, g, g, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(gh, gw), mode="bilinear")
        return torch.cat([tok, grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)

    def forward_flex(self, x, pose, nviews, taps=(8, 11)):
        """-> (stage-0 map, stage-1 map, tokens after block taps[0], tokens after block taps[1]); the token tensors
        are per view again ((b*n, 1 + h/16*w/16, C)), as forward_vit re-views them (midas/vit.py:67-71)."""
        b, _, h, w = x.shape
        gh, gw = h // self.patch_size[1], w // self.patch_size[0]
        pos = self._resize_pos_embed(self.pos_embed, gh, gw)
        bb = self.patch_embed.backbone
        s0 = bb.stages[0](bb.stem(x))
        s1 = bb.stages[1](s0)
        t = self.patch_embed.proj(bb.stages[2](s1)).flatten(2).transpose(1, 2)
        t = torch.cat((self.cls_token.expand(b, -1, -1), t), dim=1)
        t = self.pos_drop(t + pos + self.pose_embed(pose)[:, None, :])
        per_view = t.shape[1]                           # the reference hard-codes 257 = 1 + 16*16 (256x256 images)
        t = t.view(b // nviews, nviews * per_view, t.shape[2])
        kept = []
        for i, blk in enumerate(self.blocks):
            t = blk(t)
            if i in taps:
                kept.append(t.reshape(b, per_view, t.shape[2]))
        return s0, s1, kept[0], kept[1]


# ---------------------------------------------------------------------------------------------------------
# The reference's DPT wrapper (midas/vit.py:27-41,389-541; midas/blocks.py:51-85,227-342; midas/dpt_depth.py)
# ---------------------------------------------------------------------------------------------------------
class _ProjectReadout(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.project = nn.Sequential(nn.Linear(2 * dim, dim), nn.GELU())

    def forward(self, x):
        readout = x[:, 0].unsqueeze(1).expand_as(x[:, 1:])
        return self.project(torch.cat((x[:, 1:], readout), -1))


class _ResidualConvUnit(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class _FeatureFusionBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.out_conv = nn.Conv2d(c, c, 1)
        self.resConfUnit1 = _ResidualConvUnit(c)
        self.resConfUnit2 = _ResidualConvUnit(c)

    def forward(self, x, skip=None):
        if skip is not None:
            x = x + self.resConfUnit1(skip)
        x = self.resConfUnit2(x)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        return self.out_conv(x)


class _Placeholder(nn.Module):
    """Index filler so that Sequential indices (= state_dict keys) match the reference's act_postprocess stacks."""

    def forward(self, x):
        return x


class DPTHybridEncoder(nn.Module):
    """``forward(rgb, rel_cam2world_16, n_view) -> [path_2 (256 ch, H/4), path_1 (256 ch, H/2)]`` - the contract of
    ``CrossAttentionRenderer.encoder`` (models.py:178).  ``channels_last=True`` returns the maps in NHWC memory
    order, which the renderer's feature packing then takes over without a copy."""

    def __init__(self, features=256, channels_last=True):
        super().__init__()
        self.channels_last = channels_last
        self.pretrained = nn.Module()
        self.pretrained.model = VisionTransformerMultiView()
        vf = 768
        self.pretrained.act_postprocess1 = nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
        self.pretrained.act_postprocess2 = nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
        self.pretrained.act_postprocess3 = nn.Sequential(_ProjectReadout(vf), _Placeholder(), _Placeholder(),
                                                         nn.Conv2d(vf, 768, 1))
        self.pretrained.act_postprocess4 = nn.Sequential(_ProjectReadout(vf), _Placeholder(), _Placeholder(),
                                                         nn.Conv2d(vf, 768, 1), nn.Conv2d(768, 768, 3, stride=2, padding=1))
        self.scratch = nn.Module()
        for i, c in enumerate((256, 512, 768, 768), start=1):
            setattr(self.scratch, f"layer{i}_rn", nn.Conv2d(c, features, 3, padding=1, bias=False))
        for i in range(1, 5):
            setattr(self.scratch, f"refinenet{i}", _FeatureFusionBlock(features))
        # depth head of DPTDepthModel (dpt_depth.py:97-105): never evaluated (forward returns the two decoder maps),
        # present so that the reference's checkpoints load with strict=True
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, features // 2, 3, padding=1), _Placeholder(), nn.Conv2d(features // 2, 32, 3, padding=1),
            nn.ReLU(True), nn.Conv2d(32, 1, 1), nn.ReLU(True), nn.Identity())

    def _reassemble(self, post, tokens, gh, gw):
        x = post[0](tokens).transpose(1, 2)
        x = x.reshape(x.shape[0], x.shape[1], gh, gw)
        for layer in list(post)[3:]:
            x = layer(x)
        return x

    def forward(self, x, rel_transform, nviews):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        _, _, h, w = x.shape
        l1, l2, t3, t4 = self.pretrained.model.forward_flex(x, rel_transform, nviews)
        l3 = self._reassemble(self.pretrained.act_postprocess3, t3, h // 16, w // 16)
        l4 = self._reassemble(self.pretrained.act_postprocess4, t4, h // 16, w // 16)
        s = self.scratch
        p4 = s.refinenet4(s.layer4_rn(l4))
        p3 = s.refinenet3(p4, s.layer3_rn(l3))
        p2 = s.refinenet2(p3, s.layer2_rn(l2))
        p1 = s.refinenet1(p2, s.layer1_rn(l1))
        if self.channels_last:
            p2 = p2.contiguous(memory_format=torch.channels_last)
            p1 = p1.contiguous(memory_format=torch.channels_last)
        return [p2, p1]
