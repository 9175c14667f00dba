"""Drop-in ``CrossAttentionRenderer`` whose per-ray path runs in the sm_100a
CUDA library behind include/car_b200.h.

Mirrors the reference module's public surface (reference models.py:42-626):
constructor arguments and defaults (:43), parameter names / shapes
(``state_dict`` ABI, :96-145), ``get_z(input, val=False)`` (:148-188) and
``forward(input, z=None, val=False, debug=False) -> dict`` (:190-626) with the
same ``out_dict`` keys and layouts.  The 4x4 pose algebra stays in torch
(plumbing); everything from ray set-up to the final white fill is CUDA.

Not a fallback: rendering raises unless the tensors live on a CUDA device and
the library is built.
"""
import os

import torch
import torch.nn as nn

from . import _lib, packing
from .params import HOT_PATH_PARAMS


class _ResnetBlockFC(nn.Module):
    """Parameter container with the reference's names (resnet_block_fc.py:10-51)."""

    def __init__(self, size):
        super().__init__()
        self.fc_0 = nn.Linear(size, size)
        self.fc_1 = nn.Linear(size, size)
        nn.init.constant_(self.fc_0.bias, 0.0)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.constant_(self.fc_1.bias, 0.0)
        nn.init.zeros_(self.fc_1.weight)


class _ResnetFC(nn.Module):
    """Parameter container for ``phi`` (resnet_block_fc.py:65-130); evaluated in CUDA."""

    def __init__(self, d_in, n_blocks, d_out, d_latent, d_hidden):
        super().__init__()
        self.lin_in = nn.Linear(d_in, d_hidden)
        self.lin_out = nn.Linear(d_hidden, d_out)
        self.blocks = nn.ModuleList([_ResnetBlockFC(d_hidden) for _ in range(n_blocks)])
        self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(n_blocks)])
        for lin in [self.lin_in, self.lin_out, *self.lin_z]:
            nn.init.constant_(lin.bias, 0.0)
            nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")


class CrossAttentionRenderer(nn.Module):
    def __init__(self, no_sample=False, no_latent_concat=False, no_multiview=False,
                 no_high_freq=False, model="midas_vit", uv=None, repeat_attention=True, n_view=1,
                 npoints=64, num_hidden_units_phi=128, encoder=None, precision=None):
        super().__init__()
        self.n_view = n_view
        self.npoints = 64 if n_view in (1, 2) else 48          # models.py:48-51
        if npoints:
            self.npoints = npoints
        self.repeat_attention = repeat_attention
        self.no_sample = no_sample
        self.no_latent_concat = no_latent_concat
        self.no_multiview = no_multiview
        self.no_high_freq = no_high_freq
        self.model = model
        self.num_hidden_units_phi = num_hidden_units_phi
        if model != "midas_vit":
            raise NotImplementedError("only model='midas_vit' (576-channel features) is on the hot path")
        if n_view not in (1, 2, 3):
            raise NotImplementedError("the reference defines n_view in {1, 2, 3} (models.py:281,345,478)")
        if (no_sample or no_latent_concat) and n_view != 2:
            raise NotImplementedError("no_sample / no_latent_concat are built for n_view=2 (the configurations the "
                                      "reference's ablation scripts run)")
        if no_sample and no_latent_concat:
            raise NotImplementedError("no_sample and no_latent_concat together")
        if not repeat_attention:
            raise NotImplementedError("repeat_attention=False (the reference never sets it)")
        if num_hidden_units_phi != 128:
            raise NotImplementedError("num_hidden_units_phi must be 128")
        # n_view = 2 with the default flags is the fused tcgen05 hot path; every other branch runs the general
        # exact-fp32 kernels behind car_render_forward_general
        self.general = n_view != 2 or no_sample or no_latent_concat
        # image encoder: outside the hot path; plug in any module with the reference's
        # encoder.forward(rgb, cam2world_encode, n_view) -> [path_2, path_1] contract, or encoder="dpt_hybrid" for
        # the reference's own multi-view DPT-hybrid (encoder.py: 123.6 M parameters, the reference's state_dict keys).
        # The default stays None: forward(input, z=...) - the hot path - does not need one.
        if isinstance(encoder, str):
            if encoder != "dpt_hybrid":
                raise ValueError(f"unknown encoder {encoder!r} (a module, None or 'dpt_hybrid')")
            from .encoder import DPTHybridEncoder
            encoder = DPTHybridEncoder(channels_last=True)
        self.encoder = encoder
        hidden = 128
        latent = 512 + 64
        self.conv_map = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3)
        self.latent_dim = latent
        if n_view > 1 and not no_latent_concat:                        # models.py:100-105
            self.query_encode_latent = nn.Conv2d(latent + 3, latent, 1)
            self.query_encode_latent_2 = nn.Conv2d(latent, latent // 2, 1)
            self.latent_dim = latent // 2
            self.update_val_merge = nn.Conv2d(self.latent_dim * 2 + 6, self.latent_dim, 1)
        elif no_latent_concat:                                         # :106-107
            self.feature_map = nn.Conv2d(latent, latent // 2, 1)
        else:                                                          # :108
            self.update_val_merge = nn.Conv2d(latent + 6, latent, 1)
        kv_in = self.latent_dim if no_latent_concat else self.latent_dim * n_view      # :116-124
        self.latent_value = nn.Conv2d(kv_in, self.latent_dim, 1)
        self.key_map = nn.Conv2d(kv_in, hidden, 1)
        self.key_map_2 = nn.Conv2d(hidden, hidden, 1)
        self.query_embed = nn.Conv2d(16, hidden, 1)
        self.query_embed_2 = nn.Conv2d(hidden, hidden, 1)
        self.hidden_dim = hidden
        self.latent_avg_query = nn.Conv2d(9 + 16, hidden, 1)
        self.latent_avg_query_2 = nn.Conv2d(hidden, hidden, 1)
        self.latent_avg_key = nn.Conv2d(self.latent_dim, hidden, 1)
        self.latent_avg_key_2 = nn.Conv2d(hidden, hidden, 1)
        self.query_repeat_embed = nn.Conv2d(16 + 128, hidden, 1)
        self.query_repeat_embed_2 = nn.Conv2d(hidden, hidden, 1)
        self.latent_avg_repeat_query = nn.Conv2d(9 + 16 + 128, hidden, 1)
        self.latent_avg_repeat_query_2 = nn.Conv2d(hidden, hidden, 1)
        self.encode_latent = nn.Conv1d(self.latent_dim, 128, 1)
        self.phi = _ResnetFC(n_view * 9, n_blocks=3, d_out=3, d_latent=self.latent_dim * n_view,
                             d_hidden=num_hidden_units_phi)
        # B200 knobs (not part of the reference API)
        self.precision = precision or os.environ.get("CAR_PRECISION", "fp32")
        self.feature_dtype = None            # None: fp32 for fp32*, bf16 for bf16
        self.pixel_val_to_cpu = True         # reference returns pixel_val on the host (models.py:570)
        self.chunk_rays = None
        # bit 0: fused gather+encode kernel, bit 1: fused attention tail (P == 64 or 128)
        self.use_fused = int(os.environ.get("CAR_FUSED", "3"))
        self._wcache = None
        self._fcache = None
        self._ws = None
        self.last_launch_count = 0

    # ------------------------------------------------------------------ encoder side
    def get_z(self, input, val=False):
        """Feature maps [path_2 (256ch,H/4), path_1 (256ch,H/2), conv_map (64ch,H)]
        (reference models.py:148-188).  Needs an ``encoder`` module."""
        rgb = input["context"]["rgb"]
        cam2world = input["context"]["cam2world"]
        rel_cam2world = torch.matmul(torch.inverse(cam2world[:, :1]), cam2world)
        rgb = torch.flatten(rgb, 0, 1).permute(0, -1, 1, 2)
        self.H, self.W = rgb.shape[-2], rgb.shape[-1]
        rgb = (rgb + 1) / 2
        mean = rgb.new_tensor([0.485, 0.456, 0.406])[None, :, None, None]
        std = rgb.new_tensor([0.229, 0.224, 0.225])[None, :, None, None]
        rgb = (rgb - mean) / std                               # utils/util.py:21-31
        enc = rel_cam2world.reshape(-1, 16)
        if self.no_multiview:
            enc = torch.zeros_like(enc)
        if self.encoder is None:
            raise RuntimeError(
                "get_z needs an image encoder (out of the hot path, SURVEY.md §8f): pass "
                "encoder=<module> to the constructor or call forward(input, z=[z1,z2,z3])")
        if getattr(self.encoder, "channels_last", False):
            # NHWC end to end: the packed layout of the renderer is then the encoder's own output (packing.nhwc_view)
            rgb = rgb.contiguous(memory_format=torch.channels_last)
        z = list(self.encoder.forward(rgb, enc, self.n_view))
        z_conv = self.conv_map(rgb)
        if getattr(self.encoder, "channels_last", False):
            z_conv = z_conv.contiguous(memory_format=torch.channels_last)
        if self.no_high_freq:
            z_conv = torch.zeros_like(z_conv)
        return z + [z_conv]

    # ------------------------------------------------------------------ caches
    def _packed_weights(self):
        params = dict(self.named_parameters())
        key = tuple((params[n].data_ptr(), params[n]._version) for n in HOT_PATH_PARAMS)
        if self._wcache is None or self._wcache[0] != key:
            sd = {n: params[n].detach() for n in HOT_PATH_PARAMS}
            self._wcache = (key, packing.PackedWeights(sd))
        return self._wcache[1]

    def _packed_features(self, z, bf16):
        # The cache entry keeps the SOURCE tensors alive: a key made of (data_ptr, _version, shape) alone
        # would match a different scene's maps that the caching allocator placed at the recycled addresses
        # of freed ones (same shapes, _version 0), and the previous scene's packed features would be
        # rendered silently.  `is` on the held tensors cannot alias.
        c = self._fcache
        if (c is not None and c[0] == bf16 and len(c[2]) == len(z) and all(a is t_ for a, t_ in zip(c[2], z))
                and c[3] == tuple(t._version for t in z)):
            return c[1]
        self._fcache = (bf16, packing.pack_features([t.detach().float() for t in z], bf16), list(z),
                        tuple(t._version for t in z))
        return self._fcache[1]

    def release_features(self):
        """Drop the packed copy of the last scene batch (and the reference to its source maps)."""
        self._fcache = None

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    # ------------------------------------------------------------------ the hot path
    def forward(self, input, z=None, val=False, debug=False, ray_range=None, debug_taps=None):
        lib = _lib.load()
        query, context = input["query"], input["context"]
        b, n_context = context["rgb"].shape[:2]
        n_qry, R = query["uv"].shape[1:3]
        if n_context != self.n_view or n_qry != 1:
            raise NotImplementedError(f"{self.n_view} context views (n_view) and 1 query view expected, got {n_context} / {n_qry}")
        if z is None:
            z = z_orig = self.get_z(input)
        else:
            z_orig = z
        dev = z[0].device
        if dev.type != "cuda":
            raise RuntimeError("CrossAttentionRenderer (B200) has no CPU path: tensors must be on a CUDA device")
        if not hasattr(self, "H"):
            self.H, self.W = z[2].shape[-2], z[2].shape[-1]
        H, W, P = self.H, self.W, self.npoints
        if tuple(z[2].shape[-2:]) != (H, W) or z[0].shape[1] != 256 or z[2].shape[1] != 64:
            raise ValueError("z must be [ (bn,256,H/4,W/4), (bn,256,H/2,W/2), (bn,64,H,W) ]")
        prec = _lib.PRECISIONS[self.precision]
        feat_bf16 = (self.feature_dtype == "bf16") if self.feature_dtype else (prec == _lib.PREC_BF16)

        f32 = lambda t: t.to(device=dev, dtype=torch.float32)
        Cm, q = f32(context["cam2world"]), f32(query["cam2world"])
        # torch.inverse == linalg.inv_ex + a host-side check of `info` (a device sync per call, so the host
        # could never run ahead of the GPU); the same factorisation without the check keeps the call async
        inv = lambda m_: torch.linalg.inv_ex(m_, check_errors=False)[0]
        Cinv = inv(Cm)                                                  # models.py:207-208
        cams = {
            "Q": torch.matmul(Cinv, q).contiguous(),
            "Cself": torch.matmul(Cinv, Cm).contiguous(),
            "Rel": torch.stack([torch.matmul(inv(Cm[:, k:k + 1]), Cm) for k in range(n_context)],
                               dim=1).contiguous(),                     # models.py:285-286, 349-352
            "qinv": inv(q[:, 0]).contiguous(),                          # geometry.py:404
            "K": f32(context["intrinsics"]).contiguous(),
            "Kq": f32(query["intrinsics"])[:, 0].contiguous(),
        }
        uv = f32(query["uv"])[:, 0].contiguous()
        if self.no_sample:
            interval = torch.linspace(0.1, 10., P, device=dev)          # depths of the volumetric line, geometry.py:175
        else:
            interval = torch.linspace(0, 1, P, device=dev)              # models.py:261
        out = self.render_prepared(cams, uv, interval, z, b, R, ray_range=ray_range,
                                   debug_taps=debug_taps)
        out["uv"] = query["uv"]
        out["z"] = z_orig
        return out

    def render_prepared(self, cams, uv, interval, z, b, R, ray_range=None, debug_taps=None):
        """CUDA part of forward with the pose algebra already done (used directly by the
        bit-exactness tests so both sides see identical 4x4 matrices).

        With autograd enabled and a weight or feature map that requires grad, the call goes
        through ``_RenderFunction`` (training-mode forward + ``car_render_backward``): ``rgb``
        and ``depth_ray`` are then differentiable like the reference's (training.py:92,125)."""
        # The differentiable path is the TRAINING path: it runs the unfused kernels (per-sample GEMMs on tcgen05
        # with hi + lo bf16 operands, or exact fp32 when precision == "fp32_simt") on the
        # whole ray range as one chunk with every activation resident (training.py:92 calls the module in
        # train mode on 192 rays per scene).  It is therefore gated on ``self.training``: a module in
        # eval() mode always runs the configured inference precision, grad mode or not, and its outputs
        # carry no grad_fn (like calling the reference under torch.no_grad(), eval_realestate10k.py:38).
        if self.general:
            out = self._launch_general(cams, uv, interval, z, b, R, ray_range, debug_taps)
            out["at_wts"] = [out["at_wt"]]
            if self.pixel_val_to_cpu:
                out["pixel_val"] = out["pixel_val"].cpu()
            return out
        needs_grad = self.training and torch.is_grad_enabled() and debug_taps is None and (
            any(t.requires_grad for t in z)
            or any(p.requires_grad for n, p in self.named_parameters() if n in HOT_PATH_PARAMS))
        if needs_grad:
            rays = (b * R) if ray_range is None else (ray_range[1] - ray_range[0])
            need = _lib.load().car_train_workspace_bytes(self._train_precision(), self.npoints, max(1, rays))
            free = torch.cuda.mem_get_info(z[0].device)[0] + torch.cuda.memory_reserved(z[0].device) \
                - torch.cuda.memory_allocated(z[0].device)
            if need > free:
                raise RuntimeError(
                    f"training-mode forward over {rays} rays x {self.npoints} samples keeps {need / 2**30:.1f} GiB of "
                    f"activations for backward ({free / 2**30:.1f} GiB free): render fewer rays per call "
                    "(the reference trains on 192 rays per scene, train_realestate10k.py:78), or call "
                    "model.eval() / torch.no_grad() for inference")
            params = dict(self.named_parameters())
            plist = [params[n] for n in HOT_PATH_PARAMS]
            res = _RenderFunction.apply(self, cams, uv, interval, b, R, ray_range, z[0], z[1], z[2], *plist)
            out = dict(zip(_RenderFunction.OUT_KEYS, res))
        else:
            out = self._launch(cams, uv, interval, z, b, R, ray_range, debug_taps)[0]
        out["at_wts"] = [out["at_wt"]]
        if self.pixel_val_to_cpu:
            out["pixel_val"] = out["pixel_val"].cpu()                   # models.py:570
        return out

    def _train_precision(self, backward=False):
        """Arithmetic of the training forward / backward GEMMs: fp32-equivalent on the tensor cores (hi + lo bf16
        operands, three MMAs per product) unless the module was built with precision="fp32_simt" (exact fp32).
        There is no single-bf16 training mode: "bf16" modules train in the fp32-equivalent one.
        ``self.backward_precision`` ("fp32" / "fp32_simt", default: same as the forward) picks the gradient
        GEMMs separately (tests: exact forward + tensor-core backward isolates the GEMM error from ReLU flips)."""
        name = (getattr(self, "backward_precision", None) if backward else None) or self.precision
        return _lib.PREC_FP32_SIMT if name == "fp32_simt" else _lib.PREC_FP32_3XBF16

    def _launch(self, cams, uv, interval, z, b, R, ray_range=None, debug_taps=None, train=False, pw=None):
        """Fill ``car_render_args`` and enqueue ``car_render_forward``.  Returns (out, args, keep)
        where ``keep`` holds every tensor the argument struct points to."""
        lib = _lib.load()
        dev = z[0].device
        H, W, P = self.H, self.W, self.npoints
        prec = self._train_precision() if train else _lib.PRECISIONS[self.precision]
        feat_bf16 = False if train else (
            (self.feature_dtype == "bf16") if self.feature_dtype else (prec == _lib.PREC_BF16))
        if pw is None:
            pw = self._packed_weights()
        feats = self._packed_features(z, feat_bf16)
        total = b * R
        g0, g1 = (0, total) if ray_range is None else ray_range
        use_fused = 0 if train else int(self.use_fused)
        if debug_taps is not None:
            want_ = debug_taps.get("_keys")
            if want_ is None or "interp" in want_:
                use_fused &= ~3                    # interp only exists in the unfused encoder
            elif "key" in want_ or "q2" in want_ or "q1" in want_:
                use_fused &= ~2                    # key / q2 / row-major q1 only exist outside the fused tail
        if train:
            # the activations stay in this buffer until backward: it belongs to the autograd node
            ws = torch.empty(lib.car_train_workspace_bytes(prec, P, max(1, g1 - g0)), dtype=torch.uint8, device=dev)
            chunk_arg = 0
        else:
            chunk = self.chunk_rays or lib.car_default_chunk_rays(prec, P, use_fused)
            chunk = max(1, min(chunk, g1 - g0))
            ws = self._workspace(lib.car_workspace_bytes(prec, P, chunk, use_fused), dev)
            chunk_arg = chunk
        out = {
            "rgb": torch.zeros(b, 1, R, 3, device=dev),
            "valid_mask": torch.zeros(b, R, 1, device=dev),
            "depth_ray": torch.zeros(b, R, 1, device=dev),
            "at_wt": torch.zeros(b * 2, R, P, device=dev),
            "at_wt_max": torch.zeros(b * 2, R, 1, dtype=torch.int64, device=dev),
            "pixel_val": torch.zeros(b * 2, R, P, 2, device=dev),
            "coords": torch.zeros(b * 2, R, 9, device=dev),
        }
        a = _lib.car_render_args()
        a.abi_version = _lib.ABI_VERSION
        a.precision = prec
        a.b, a.R, a.P, a.H, a.W = b, R, P, H, W
        a.ray_begin, a.ray_end = g0, g1
        a.feat_bf16 = int(feat_bf16)
        for i in range(3):
            a.feat[i] = feats[i].data_ptr()
        a.weights = pw.c_struct()
        for k in ("Q", "Cself", "Rel", "qinv", "K", "Kq"):
            assert cams[k].is_contiguous() and cams[k].dtype == torch.float32 and cams[k].device == dev
            setattr(a.cams, k, cams[k].data_ptr())
        a.uv, a.interval = uv.data_ptr(), interval.data_ptr()
        for k, t in out.items():
            setattr(a, k, t.data_ptr())
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        if debug_taps is not None:
            rows = (g1 - g0) * 2 * P
            shapes = {"geom": (rows, _lib.GEOM_STRIDE), "x": (rows, 2, _lib.K_ENC), "interp": (rows, 576),
                      "value": (rows, 288), "key": (rows, 128), "q1": (rows, 128), "q2": (rows, 128),
                      "zfinal": (g1 - g0, 288)}
            want = debug_taps.pop("_keys", None)
            for k, shp in shapes.items():
                if k == "x" and prec != _lib.PREC_FP32_SIMT:
                    continue
                if want is not None and k not in want:
                    continue
                debug_taps[k] = torch.zeros(*shp, device=dev)
                setattr(a.debug, k, debug_taps[k].data_ptr())
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        a.use_fused = use_fused
        a.train = int(train)
        a.chunk_rays = chunk_arg
        with torch.cuda.device(dev):
            _lib.check(lib.car_render_forward(a), "car_render_forward")
        self.last_launch_count = lib.car_last_launch_count()
        keep = (pw, feats, ws, cams, uv, interval)
        return out, a, keep


def _launch_general(self, cams, uv, interval, z, b, R, ray_range=None, debug_taps=None):
    """Fill ``car_general_args`` and enqueue ``car_render_forward_general`` (n_view 1 / 3, no_sample,
    no_latent_concat; inference only: these branches have no backward kernels)."""
    lib = _lib.load()
    dev = z[0].device
    H, W, P, n = self.H, self.W, self.npoints, self.n_view
    flags = (_lib.FLAG_NO_SAMPLE if self.no_sample else 0) | (_lib.FLAG_NO_LATENT_CONCAT if self.no_latent_concat else 0)
    params = {k: v for k, v in self.named_parameters() if not k.startswith("encoder.")}
    key = tuple((k, v.data_ptr(), v._version) for k, v in params.items())
    if self._wcache is None or self._wcache[0] != key:
        self._wcache = (key, packing.PackedGeneralWeights({k: v.detach() for k, v in params.items()}, n,
                                                          no_latent_concat=self.no_latent_concat))
    pw = self._wcache[1]
    feats = self._packed_features(z, False)
    total = b * R
    g0, g1 = (0, total) if ray_range is None else ray_range
    chunk = self.chunk_rays or lib.car_general_default_chunk_rays(n, flags, P)
    chunk = max(1, min(chunk, g1 - g0))
    ws = self._workspace(lib.car_general_workspace_bytes(n, flags, P, chunk), dev)
    out = {
        "rgb": torch.zeros(b, 1, R, 3, device=dev),
        "valid_mask": torch.zeros(b, R, 1, device=dev),
        "depth_ray": torch.zeros(b, R, 1, device=dev),
        "at_wt": torch.zeros(b * n, R, P, device=dev),
        "at_wt_max": torch.zeros(b * n, R, 1, dtype=torch.int64, device=dev),
        "pixel_val": torch.zeros(b * n, R, P, 2, device=dev),
        "coords": torch.zeros(b * n, R, 9, device=dev),
    }
    a = _lib.car_general_args()
    a.abi_version = _lib.ABI_VERSION
    a.n_view, a.flags = n, flags
    a.b, a.R, a.P, a.H, a.W = b, R, P, H, W
    a.ray_begin, a.ray_end = g0, g1
    for i in range(3):
        a.feat[i] = feats[i].data_ptr()
    a.weights = pw.c_struct()
    want = {"Q": (b, n, 4, 4), "Cself": (b, n, 4, 4), "Rel": (b, n, n, 4, 4), "qinv": (b, 4, 4), "K": (b, n, 4, 4), "Kq": (b, 4, 4)}
    for k, shp in want.items():
        assert cams[k].is_contiguous() and cams[k].dtype == torch.float32 and cams[k].device == dev and tuple(cams[k].shape) == shp, k
        setattr(a.cams, k, cams[k].data_ptr())
    a.uv, a.interval = uv.data_ptr(), interval.data_ptr()
    for k, t in out.items():
        setattr(a, k, t.data_ptr())
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.stream = torch.cuda.current_stream(dev).cuda_stream
    a.chunk_rays = chunk
    # "fp32": per-sample GEMMs on tcgen05 with hi + lo bf16 operands; "fp32_simt": exact fp32; "bf16" has no
    # single-MMA kernels on these branches and runs the (more accurate) hi + lo ones
    a.precision = _lib.PREC_FP32_SIMT if self.precision == "fp32_simt" else _lib.PRECISIONS["fp32"]
    if debug_taps is not None:
        rows = (g1 - g0) * n * P
        ci = 576 if (self.no_latent_concat or n == 1) else 288 * n
        debug_taps["interp"] = torch.zeros(rows, ci, device=dev)
        debug_taps["zfinal"] = torch.zeros(g1 - g0, self.latent_dim, device=dev)
        a.debug_interp, a.debug_zfinal = debug_taps["interp"].data_ptr(), debug_taps["zfinal"].data_ptr()
    with torch.cuda.device(dev):
        _lib.check(lib.car_render_forward_general(a), "car_render_forward_general")
    self.last_launch_count = lib.car_last_launch_count()
    return out


CrossAttentionRenderer._launch_general = _launch_general


class _RenderFunction(torch.autograd.Function):
    """Autograd node of the CUDA path: ``car_render_forward(train=1)`` /
    ``car_render_backward`` (reference: autograd through models.py:278-621).  Differentiable
    inputs are the three feature maps and the renderer weights; differentiable outputs are
    ``rgb`` and ``depth_ray`` (the only ones the reference's losses read,
    loss_functions.py:74-123)."""

    OUT_KEYS = ("rgb", "depth_ray", "valid_mask", "at_wt", "at_wt_max", "pixel_val", "coords")

    @staticmethod
    def forward(ctx, model, cams, uv, interval, b, R, ray_range, z1, z2, z3, *params):
        sd = {n: p.detach() for n, p in zip(HOT_PATH_PARAMS, params)}
        pw = packing.PackedWeights(sd, folds=False)
        z = [z1.detach(), z2.detach(), z3.detach()]
        out, a, keep = model._launch(cams, uv, interval, z, b, R, ray_range, train=True, pw=pw)
        ctx.args, ctx.keep, ctx.pw = a, keep, pw
        ctx.backward_precision = model._train_precision(backward=True)
        ctx.z_like = z
        ctx.param_shapes = {n: p.shape for n, p in zip(HOT_PATH_PARAMS, params)}
        res = tuple(out[k] for k in _RenderFunction.OUT_KEYS)
        ctx.mark_non_differentiable(*res[2:])
        ctx.save_for_backward(out["at_wt"])          # read by the round-1 attention backward
        return res

    @staticmethod
    def backward(ctx, d_rgb, d_depth, *_):
        lib = _lib.load()
        a, pw = ctx.args, ctx.pw
        (at_wt,) = ctx.saved_tensors
        assert at_wt.data_ptr() == a.at_wt
        dev = ctx.z_like[0].device
        nr = a.ray_end - a.ray_begin
        grads = packing.PackedGrads(pw)
        want_feat = any(ctx.needs_input_grad[7:10])
        d_feat = [torch.zeros(t.shape[0], t.shape[2], t.shape[3], t.shape[1], device=dev) for t in ctx.z_like] \
            if want_feat else None
        bprec = ctx.backward_precision
        ws = torch.empty(lib.car_backward_workspace_bytes(bprec, a.P, nr), dtype=torch.uint8, device=dev)
        bw = _lib.car_backward_args()
        bw.abi_version = _lib.ABI_VERSION
        bw.fwd = C_pointer(a)
        if d_rgb is not None:
            d_rgb = d_rgb.contiguous().float()
            bw.d_rgb = d_rgb.data_ptr()
        if d_depth is not None:
            d_depth = d_depth.contiguous().float()
            bw.d_depth_ray = d_depth.data_ptr()
        bw.grads = grads.c_struct()
        if want_feat:
            for i in range(3):
                bw.d_feat[i] = d_feat[i].data_ptr()
        bw.workspace, bw.workspace_bytes = ws.data_ptr(), ws.numel()
        bw.stream = torch.cuda.current_stream(dev).cuda_stream
        bw.precision = bprec
        with torch.cuda.device(dev):
            _lib.check(lib.car_render_backward(bw), "car_render_backward")
        gsd = grads.unpack(ctx.param_shapes)
        gz = packing.unpack_feature_grads(d_feat, ctx.z_like) if want_feat else [None, None, None]
        pgr = tuple(gsd[n] if ctx.needs_input_grad[10 + i] else None for i, n in enumerate(HOT_PATH_PARAMS))
        return (None,) * 7 + tuple(gz) + pgr


def C_pointer(struct):
    import ctypes
    return ctypes.pointer(struct)
