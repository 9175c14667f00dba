"""Weight packing: reference ``state_dict`` tensors -> the K-major matrices of
``car_weights`` (include/car_b200.h).

Pure re-layout done once per weight update with torch ops on the device
(plumbing): every matrix becomes ``[N][K]`` fp32 with K zero-padded, plus its
bf16 ``hi`` and ``lo = bf16(w - hi)`` halves for the tensor-core precisions.

Two algebraic re-groupings (exact in real arithmetic, ~1 ulp in fp32):
  * ``query_repeat_embed`` (144 -> 128, reference models.py:136,552-553) is
    split into its per-ray block (columns 0..127, applied to
    ``encode_latent(z_local)``) and its per-sample block (columns 128..143,
    applied to ``local_coords``);
  * ``phi.lin_z[i]`` (576 -> 128) sees the same 288-vector twice
    (models.py:604-606), so its two column halves are summed.
"""
import torch

from . import _lib


def _split_bf16(w):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def _aligned(t):
    """The kernels read weights with 16-byte vector loads / TMA: a view into some larger buffer (e.g. a flat
    optimiser buffer) at an odd offset is copied to a fresh allocation."""
    return t if t.data_ptr() % 16 == 0 else t.clone()


class PackedMat:
    def __init__(self, w, bias, k_pad=None):
        w = w.detach().float().reshape(w.shape[0], -1)
        if k_pad is not None and k_pad != w.shape[1]:
            w = torch.nn.functional.pad(w, (0, k_pad - w.shape[1]))
        self.f32 = _aligned(w.contiguous())
        self.hi, self.lo = _split_bf16(self.f32)
        self.bias = None if bias is None else _aligned(bias.detach().float().contiguous())
        self.N, self.K = self.f32.shape

    def c_struct(self):
        m = _lib.car_mat()
        m.f32 = self.f32.data_ptr()
        m.hi = self.hi.data_ptr()
        m.lo = self.lo.data_ptr()
        m.bias = self.bias.data_ptr() if self.bias is not None else None
        m.N, m.K = self.N, self.K
        return m


_FOLD64_PERM = None


def _fold64_perm():
    """K-column order of ``kv_fold64``: the fused kernel's 64-wide H stages (include/car_b200.h)."""
    global _FOLD64_PERM
    if _FOLD64_PERM is None:
        perm = torch.empty(1152, dtype=torch.long)
        for v in range(2):
            for q in range(9):
                c, j = divmod(q, 3)
                for h in range(2):
                    dst = v * 576 + q * 64 + h * 32
                    src = v * 576 + c * 192 + h * 96 + j * 32
                    perm[dst:dst + 32] = torch.arange(src, src + 32)
        _FOLD64_PERM = perm
    return _FOLD64_PERM


class PackedWeights:
    """Holds the packed device tensors alive and exposes the ctypes struct.  ``folds=False`` (training: the unfused
    kernels read the plain layers only) skips the composed matrices of the fused inference kernels."""

    def __init__(self, sd, folds=True):
        g = lambda n: sd[n]
        P = PackedMat
        self.m = {}
        self.m["enc1"] = P(g("query_encode_latent.weight"), g("query_encode_latent.bias"), _lib.K_ENC)
        self.m["enc2"] = P(g("query_encode_latent_2.weight"), g("query_encode_latent_2.bias"))
        self.m["value"] = P(g("latent_value.weight"), g("latent_value.bias"))
        self.m["key1"] = P(g("key_map.weight"), g("key_map.bias"))
        self.m["key2"] = P(g("key_map_2.weight"), g("key_map_2.bias"))
        self.m["qry1"] = P(g("query_embed.weight"), g("query_embed.bias"))
        self.m["qry2"] = P(g("query_embed_2.weight"), g("query_embed_2.bias"))
        rep = g("query_repeat_embed.weight").reshape(128, 144)
        self.m["rep1_g"] = P(rep[:, :128], g("query_repeat_embed.bias"))
        self.m["rep1_loc"] = P(rep[:, 128:], None)
        self.m["rep2"] = P(g("query_repeat_embed_2.weight"), g("query_repeat_embed_2.bias"))
        self.m["enc_lat"] = P(g("encode_latent.weight"), g("encode_latent.bias"))
        self.m["phi_in"] = P(g("phi.lin_in.weight"), g("phi.lin_in.bias"), 32)
        for i in range(3):
            wz = g(f"phi.lin_z.{i}.weight")
            self.m[f"phi_z{i}"] = P(wz[:, :288] + wz[:, 288:], g(f"phi.lin_z.{i}.bias"))
            self.m[f"phi_fc0{i}"] = P(g(f"phi.blocks.{i}.fc_0.weight"), g(f"phi.blocks.{i}.fc_0.bias"))
            self.m[f"phi_fc1{i}"] = P(g(f"phi.blocks.{i}.fc_1.weight"), g(f"phi.blocks.{i}.fc_1.bias"))
        self.m["phi_out"] = P(g("phi.lin_out.weight"), g("phi.lin_out.bias"))
        if not folds:
            return
        # fold query_encode_latent_2 into [latent_value ; key_map] (float64 product, rounded once)
        w2 = g("query_encode_latent_2.weight").reshape(288, 576).double()
        b2 = g("query_encode_latent_2.bias").double()
        w3 = torch.cat([g("latent_value.weight").reshape(288, 576), g("key_map.weight").reshape(128, 576)], 0).double()
        b3 = torch.cat([g("latent_value.bias"), g("key_map.bias")], 0).double()
        fold = torch.cat([w3[:, :288] @ w2, w3[:, 288:] @ w2], dim=1)            # (416, 1152)
        bfold = w3 @ torch.cat([b2, b2]) + b3
        self.m["kv_fold"] = P(fold.float(), bfold.float())
        # the same matrix with its K columns in the order of the fused kernel's 64-wide H stages
        # (include/car_b200.h, car_weights::kv_fold64)
        self.m["kv_fold64"] = P(fold.float()[:, _fold64_perm().to(fold.device)], bfold.float())
        # colour MLP: all ten matrices along K in 64-column blocks (include/car_b200.h, car_weights::phi_pack)
        pad = lambda w_, k_: torch.nn.functional.pad(w_, (0, k_ - w_.shape[1]))
        parts = [pad(self.m["phi_in"].f32, 64)]
        for i in range(3):
            parts += [pad(self.m[f"phi_z{i}"].f32, 320), self.m[f"phi_fc0{i}"].f32, self.m[f"phi_fc1{i}"].f32]
        self.m["phi_pack"] = P(torch.cat(parts, dim=1), None)
        # per-ray row bias of the round-2 query MLP: query_repeat_embed[:, :128] o encode_latent (float64 product)
        wg = rep[:, :128].double()
        wel = g("encode_latent.weight").reshape(128, 288).double()
        self.m["rowb_fold"] = P((wg @ wel).float(),
                                (wg @ g("encode_latent.bias").double() + g("query_repeat_embed.bias").double()).float())

    def c_struct(self):
        w = _lib.car_weights()
        for name in ("enc1", "enc2", "value", "key1", "key2", "qry1", "qry2", "rep1_loc",
                     "rep1_g", "rep2", "enc_lat", "phi_in", "phi_out", "kv_fold", "kv_fold64", "rowb_fold", "phi_pack"):
            if name in self.m:                       # folds=False leaves the composed matrices null
                setattr(w, name, self.m[name].c_struct())
        for i in range(3):
            w.phi_z[i] = self.m[f"phi_z{i}"].c_struct()
            w.phi_fc0[i] = self.m[f"phi_fc0{i}"].c_struct()
            w.phi_fc1[i] = self.m[f"phi_fc1{i}"].c_struct()
        return w


class PackedGeneralWeights:
    """``car_general_weights`` (include/car_b200.h) for the other forward branches: n_view in {1, 3} and the
    no_latent_concat ablation (no_sample uses the n_view = 2 weights unchanged).  Re-groupings on top of
    ``PackedWeights``': for n_view = 3 the columns of ``latent_value`` / ``key_map`` are re-ordered from the
    reference's channel-interleaved order (index 3c + k, models.py:444-446) to part-major (k*288 + c), the
    order in which the kernels write the three encoded parts of a row; ``phi.lin_z`` is summed over its
    n_view identical column blocks (models.py:604-606); ``phi.lin_in`` is padded from 9*n_view to 32."""

    def __init__(self, sd, n_view, no_latent_concat=False):
        g = lambda n: sd[n]
        P = PackedMat
        self.m = {}
        concat = n_view > 1 and not no_latent_concat
        L = 288 if concat else 576
        if concat:
            self.m["enc1"] = P(g("query_encode_latent.weight"), g("query_encode_latent.bias"), _lib.K_ENC)
            self.m["enc2"] = P(g("query_encode_latent_2.weight"), g("query_encode_latent_2.bias"))
        elif n_view == 1:
            self.m["merge"] = P(g("update_val_merge.weight"), g("update_val_merge.bias"), _lib.K_ENC)
        wv = g("latent_value.weight").reshape(L, -1)
        wk = g("key_map.weight").reshape(128, -1)
        if concat and n_view == 3:
            perm = torch.tensor([3 * c + k for k in range(3) for c in range(288)], device=wv.device)
            wv, wk = wv[:, perm], wk[:, perm]
        self.m["value"] = P(wv, g("latent_value.bias"))
        self.m["key1"] = P(wk, g("key_map.bias"))
        self.m["key2"] = P(g("key_map_2.weight"), g("key_map_2.bias"))
        self.m["qry1"] = P(g("query_embed.weight"), g("query_embed.bias"))
        self.m["qry2"] = P(g("query_embed_2.weight"), g("query_embed_2.bias"))
        rep = g("query_repeat_embed.weight").reshape(128, 144)
        self.m["rep1_g"] = P(rep[:, :128], g("query_repeat_embed.bias"))
        self.m["rep1_loc"] = P(rep[:, 128:], None)
        self.m["rep2"] = P(g("query_repeat_embed_2.weight"), g("query_repeat_embed_2.bias"))
        self.m["enc_lat"] = P(g("encode_latent.weight"), g("encode_latent.bias"))
        self.m["phi_in"] = P(g("phi.lin_in.weight"), g("phi.lin_in.bias"), 32)
        for i in range(3):
            wz = g(f"phi.lin_z.{i}.weight")
            self.m[f"phi_z{i}"] = P(sum(wz[:, q * L:(q + 1) * L] for q in range(n_view)), g(f"phi.lin_z.{i}.bias"))
            self.m[f"phi_fc0{i}"] = P(g(f"phi.blocks.{i}.fc_0.weight"), g(f"phi.blocks.{i}.fc_0.bias"))
            self.m[f"phi_fc1{i}"] = P(g(f"phi.blocks.{i}.fc_1.weight"), g(f"phi.blocks.{i}.fc_1.bias"))
        self.m["phi_out"] = P(g("phi.lin_out.weight"), g("phi.lin_out.bias"))

    def c_struct(self):
        w = _lib.car_general_weights()
        for name in ("enc1", "enc2", "merge", "value", "key1", "key2", "qry1", "qry2", "rep1_loc", "rep1_g", "rep2",
                     "enc_lat", "phi_in", "phi_out"):
            if name in self.m:
                setattr(w, name, self.m[name].c_struct())
        for i in range(3):
            w.phi_z[i] = self.m[f"phi_z{i}"].c_struct()
            w.phi_fc0[i] = self.m[f"phi_fc0{i}"].c_struct()
            w.phi_fc1[i] = self.m[f"phi_fc1{i}"].c_struct()
        return w


class PackedGrads:
    """Zero-initialised gradient buffers in the packed layout of ``car_weights``
    (``car_weight_grads``), and the map back to ``state_dict`` shapes."""

    def __init__(self, pw):
        self.g = {}
        for name, m in pw.m.items():
            if name in ("kv_fold", "kv_fold64", "rowb_fold", "phi_pack"):
                continue
            w = torch.zeros_like(m.f32)
            b = None if m.bias is None else torch.zeros_like(m.bias)
            self.g[name] = (w, b)

    def c_struct(self):
        s = _lib.car_weight_grads()

        def fill(dst, name):
            w, b = self.g[name]
            dst.w = w.data_ptr()
            dst.bias = b.data_ptr() if b is not None else None
        for name in _lib.GRAD_MATS + ("phi_out",):
            fill(getattr(s, name), name)
        for i in range(3):
            fill(s.phi_z[i], f"phi_z{i}")
            fill(s.phi_fc0[i], f"phi_fc0{i}")
            fill(s.phi_fc1[i], f"phi_fc1{i}")
        return s

    def unpack(self, shapes):
        """-> {state_dict name: gradient tensor of that parameter's shape}.  Inverse of the
        re-groupings PackedWeights applies (K padding, query_repeat_embed split, lin_z fold)."""
        g = self.g
        out = {}

        def put(name, w, b):
            out[name + ".weight"] = w.reshape(shapes[name + ".weight"])
            if b is not None:
                out[name + ".bias"] = b
        put("query_encode_latent", g["enc1"][0][:, :579].contiguous(), g["enc1"][1])
        put("query_encode_latent_2", *g["enc2"])
        put("latent_value", *g["value"])
        put("key_map", *g["key1"])
        put("key_map_2", *g["key2"])
        put("query_embed", *g["qry1"])
        put("query_embed_2", *g["qry2"])
        put("query_repeat_embed", torch.cat([g["rep1_g"][0], g["rep1_loc"][0]], dim=1), g["rep1_g"][1])
        put("query_repeat_embed_2", *g["rep2"])
        put("encode_latent", *g["enc_lat"])
        put("phi.lin_in", g["phi_in"][0][:, :18].contiguous(), g["phi_in"][1])
        for i in range(3):
            wz = g[f"phi_z{i}"][0]
            put(f"phi.lin_z.{i}", torch.cat([wz, wz], dim=1), g[f"phi_z{i}"][1])
            put(f"phi.blocks.{i}.fc_0", *g[f"phi_fc0{i}"])
            put(f"phi.blocks.{i}.fc_1", *g[f"phi_fc1{i}"])
        put("phi.lin_out", *g["phi_out"])
        return out


def unpack_feature_grads(packed, like):
    """Packed NHWC fp32 gradient buffers -> NCHW tensors shaped like the encoder maps."""
    lib = _lib.load()
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    for o, t in zip(packed, like):
        bn, Cc, h, w = t.shape
        d = torch.empty(bn, Cc, h, w, dtype=torch.float32, device=o.device)
        _lib.check(lib.car_unpack_features(o.data_ptr(), d.data_ptr(), bn, Cc, h, w, stream),
                   "car_unpack_features")
        out.append(d)
    return out


def nhwc_view(t):
    """The NHWC buffer of a map that already lies in channels_last memory order (what
    ``encoder.DPTHybridEncoder(channels_last=True)`` and cuDNN's channels_last convolutions produce), as a
    contiguous (bn, h, w, C) view - or None when the map is NCHW and has to go through ``car_pack_features``."""
    if t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last):
        v = t.permute(0, 2, 3, 1)
        assert v.is_contiguous()
        return v
    return None


def pack_features(z, bf16=False):
    """[z1,z2,z3] NCHW fp32 (device) -> list of packed NHWC buffers via
    car_pack_features.  One-off per scene batch.  Maps that are already channels_last are taken over as they are
    (fp32: zero-copy view; bf16: one conversion pass)."""
    lib = _lib.load()
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    for t in z:
        assert t.is_cuda and t.dtype == torch.float32, "features must be fp32 CUDA tensors"
        v = nhwc_view(t)
        if v is not None:
            out.append(v.to(torch.bfloat16) if bf16 else v)
            continue
        t = t.contiguous()
        bn, Cc, h, w = t.shape
        o = torch.empty(bn, h, w, Cc, dtype=torch.bfloat16 if bf16 else torch.float32, device=t.device)
        _lib.check(lib.car_pack_features(t.data_ptr(), o.data_ptr(), bn, Cc, h, w, int(bf16), stream),
                   "car_pack_features")
        out.append(o)
    return out
