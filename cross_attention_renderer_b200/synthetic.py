"""Deterministic synthetic scenes, feature maps and weights.

Everything is generated on a CPU ``torch.Generator`` so that the same bytes
are produced in this container (golden generation against the reference) and
on the GPU box (parity tests, bench).  No file in the reference is read here.

Input dict layout follows the dataset contract the reference consumes
(reference dataset/realestate10k_dataio.py:456-466, models.py:195-198):
``context.{rgb (b,n,H,W,3), cam2world (b,n,4,4), intrinsics (b,n,4,4)}``,
``query.{cam2world (b,1,4,4), intrinsics (b,1,4,4), uv (b,1,R,2)}``.
"""
import math

import torch

from .params import renderer_param_shapes


def _rot_y(theta):
    c, s = math.cos(theta), math.sin(theta)
    m = torch.eye(4, dtype=torch.float64)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def _trans(x, y, z):
    m = torch.eye(4, dtype=torch.float64)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def _rot_axis(axis, theta):
    """Rodrigues rotation about ``axis`` (3,) by ``theta`` as a 4x4."""
    a = axis / axis.norm().clamp_min(1e-12)
    Kx = torch.zeros(3, 3, dtype=torch.float64)
    Kx[0, 1], Kx[0, 2], Kx[1, 0], Kx[1, 2], Kx[2, 0], Kx[2, 1] = -a[2], a[1], a[2], -a[0], -a[1], a[0]
    R = torch.eye(3, dtype=torch.float64) + math.sin(theta) * Kx + (1 - math.cos(theta)) * (Kx @ Kx)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, :3] = R
    return m


def _look_at(pos, target):
    """cam2world of a camera at ``pos`` whose +z axis points at ``target`` (y roughly down-world)."""
    zax = target - pos
    zax = zax / zax.norm()
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    xax = torch.linalg.cross(up, zax)
    xax = xax / xax.norm()
    yax = torch.linalg.cross(zax, xax)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = xax, yax, zax, pos
    return m


def make_intrinsics(H):
    """Pinhole f=225 px at 256² (reference dataset/load_video_superglue.py:465)."""
    k = torch.eye(4, dtype=torch.float32)
    k[0, 0] = k[1, 1] = 225.0 * H / 256.0
    k[0, 2] = k[1, 2] = H / 2.0
    return k


def make_uv(Ht, Wt):
    """Full target grid, row-major, (x, y) pixel floats (realestate10k_dataio.py:238-245)."""
    ys, xs = torch.meshgrid(torch.arange(Ht, dtype=torch.float32),
                            torch.arange(Wt, dtype=torch.float32), indexing="ij")
    return torch.stack([xs, ys], dim=-1).reshape(-1, 2)


def make_inputs(b, H, Ht=None, Wt=None, seed=0, mode="default", rays=None, n_ctx=2):
    """Build the reference-style ``input`` dict for ``b`` scenes, ``n_ctx`` context views (2 unless
    stated; 1 keeps the first camera of the pair, 3 adds one below the baseline).

    mode: "sweep"    bit-exactness sweep: per scene a random query pose (rotation up to
                     ~0.6 rad about a random axis, translation N(0, 0.5)), and every scene with s % 6 != 0 one
                     degenerate configuration: 1 the central target ray passes through context camera 0's
                     centre, 2 the query looks along context 0's x axis (central ray parallel to its image
                     plane), 3 the query sits behind context camera 0 looking forward, 4 query at the context
                     camera, 5 query far outside both frusta;
          "default"  wide-baseline converging pair, query between them;
          "outside"  query far outside both frusta looking away (white rays,
                     all-invalid tie-break of epipolar.py:142);
          "mixed"    per-scene alternation of the two plus a query that sits
                     exactly at context camera 0 (origin-at-camera branch,
                     epipolar.py:213-215).
    """
    Ht = Ht or H
    Wt = Wt or Ht
    g = torch.Generator().manual_seed(1000 + seed)
    K = make_intrinsics(H)
    # query intrinsics are expressed at the target resolution
    Kq = make_intrinsics(Ht)
    ctx_c2w, qry_c2w = [], []
    for s in range(b):
        jit = torch.randn(max(2, n_ctx), 3, generator=g, dtype=torch.float64) * 0.02
        c0 = _trans(-0.3 + jit[0, 0], jit[0, 1], jit[0, 2]) @ _rot_y(+0.1)
        c1 = _trans(+0.3 + jit[1, 0], jit[1, 1], jit[1, 2]) @ _rot_y(-0.1)
        c2 = _trans(jit[2, 0], -0.25 + jit[2, 1], 0.05 + jit[2, 2]) @ _rot_y(0.02) if n_ctx == 3 else None
        m = mode
        if mode == "mixed":
            m = ("default", "outside", "at_camera")[s % 3]
        if mode == "sweep":
            m = ("random", "through_centre", "parallel", "behind", "at_camera", "outside")[s % 6]
        if m == "random":
            ax = torch.randn(3, generator=g, dtype=torch.float64)
            q = _trans(*(torch.randn(3, generator=g, dtype=torch.float64) * 0.5).tolist()) @ _rot_axis(ax, 0.6 * float(torch.rand(1, generator=g, dtype=torch.float64)))
        elif m == "through_centre":
            # query placed in front of context camera 0, looking straight at its centre
            pos = c0 @ torch.tensor([0.4, -0.2, 1.5, 1.0], dtype=torch.float64)
            q = _look_at(pos[:3], c0[:3, 3])
        elif m == "parallel":
            q = c0.clone() @ _trans(0.0, 0.0, 1.0) @ _rot_y(math.pi / 2)
        elif m == "behind":
            q = c0.clone() @ _trans(0.05, -0.03, -1.0)
        elif m == "default":
            q = _trans(0.05 * (s % 5) - 0.1, 0.02 * (s % 3), 0.03 * (s % 2))
        elif m == "outside":
            q = _trans(5.0, 0.3, -2.0) @ _rot_y(2.6)
        elif m == "at_camera":
            q = c0.clone() @ _rot_y(0.05)
        else:
            raise ValueError(mode)
        ctx_c2w.append(torch.stack([c0, c1, c2][:n_ctx]))
        qry_c2w.append(q[None])
    ctx_c2w = torch.stack(ctx_c2w).float()
    qry_c2w = torch.stack(qry_c2w).float()
    uv = make_uv(Ht, Wt)
    if rays is not None:
        idx = torch.randperm(uv.shape[0], generator=g)[:rays].sort().values
        uv = uv[idx]
    uv = uv[None, None].expand(b, 1, -1, -1).contiguous()
    return {
        "context": {
            "rgb": torch.zeros(b, n_ctx, H, H, 3),
            "cam2world": ctx_c2w,
            "intrinsics": K[None, None].expand(b, n_ctx, -1, -1).contiguous(),
        },
        "query": {
            "cam2world": qry_c2w,
            "intrinsics": Kq[None, None].expand(b, 1, -1, -1).contiguous(),
            "uv": uv,
        },
    }


def make_features(b, H, seed=0, n_view=2, dtype=torch.float32):
    """z = [z1 (b·n,256,H/4,H/4), z2 (b·n,256,H/2,H/2), z3 (b·n,64,H,H)] ~ N(0,1)
    (layout of get_z's return for midas_vit, models.py:178-188)."""
    g = torch.Generator().manual_seed(2000 + seed)
    bn = b * n_view
    z1 = torch.randn(bn, 256, H // 4, H // 4, generator=g)
    z2 = torch.randn(bn, 256, H // 2, H // 2, generator=g)
    z3 = torch.randn(bn, 64, H, H, generator=g)
    return [z1.to(dtype), z2.to(dtype), z3.to(dtype)]


def make_state_dict(seed=0, peaky=False, n_view=2, no_latent_concat=False):
    """Every renderer parameter re-randomised: W ~ N(0, 1/sqrt(fan_in)),
    b ~ N(0, 0.1).  (The reference's default init zeroes phi.blocks.*.fc_1,
    resnet_block_fc.py:39, which would hide bugs in those layers.)
    ``peaky`` scales the last key/query layers ×8 so the softmax is far from
    uniform."""
    g = torch.Generator().manual_seed(3000 + seed)
    sd = {}
    for name, shape in renderer_param_shapes(n_view, no_latent_concat=no_latent_concat).items():
        if name.endswith(".weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(*shape, generator=g) / math.sqrt(fan_in)
            if peaky and name.split(".")[0] in ("key_map_2", "query_embed_2"):
                w = w * 8.0
            sd[name] = w
        else:
            sd[name] = torch.randn(*shape, generator=g) * 0.1
    return sd


def to_device(obj, device):
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, device) for v in obj)
    if torch.is_tensor(obj):
        return obj.to(device)
    return obj
