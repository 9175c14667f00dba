"""ORACLE — test infrastructure, NOT product code.

CPU (or any torch device) restatement of the reference's per-ray rendering
path ``CrossAttentionRenderer.forward(input, z=z)`` (reference models.py:190-626):
``render`` is the n_view=2 hot path (plus its ``no_sample`` / ``no_latent_concat``
ablations), ``render_single_view`` / ``render_three_views`` the other branches - those
and the ablations are oracle-only so far (tests/test_oracle_nview.py).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this file; the product
(``cross_attention_renderer_b200``) never does.

Parity pinning: ``tests/test_oracle_golden.py`` checks this restatement
against golden vectors produced by executing the unmodified reference in the
build container (``tests/golden/make_golden.py``): integer outputs (bilinear
tap indices, ``valid_mask``, ``at_wt_max``) exactly, floats to a few ulp at
the geometry stages and 1e-5 relative downstream.

Unlike the reference, every stage with a "bit-exact" claim (A.1-A.3: ray
set-up, epipolar clipping, line samples, and the integer bilinear taps
derived from them) is written as scalar IEEE-754 fp32 operations in a fixed
order with no fused multiply-add and no library matmul, so that a CUDA kernel
using separately rounded mul/add/div/sqrt reproduces it bit for bit.  Those
stages run on the ``X32`` wrapper below, which evaluates every fp32 operation
in float64 and rounds once to fp32 (exactly the IEEE result for + - * / sqrt):
torch's own fp32 kernels are NOT a usable definition, because they differ
between devices (measured on torch 2.11: CPU ``sqrt`` is not correctly rounded
- 0.65 % of random inputs differ from ``sqrt.rn`` -, CUDA ``x / python_scalar``
multiplies by the reciprocal; scripts/diag_ops.py).  Where the reference calls
a matmul/bmm/norm whose internal order is a library detail (geometry.py:417,
epipolar.py:25, F.normalize) the order chosen here is documented inline.

Stage names A.0 ... A.12 follow SURVEY.md Appendix A.
"""
import math

import torch
import torch.nn.functional as F

INF = float("inf")


class X32:
    """fp32 tensor whose arithmetic is exactly-rounded IEEE binary32, evaluated through
    float64 (double rounding is innocuous for + - * / sqrt since 53 >= 2*24 + 2).
    Python scalars are first rounded to fp32, like literals with an ``f`` suffix."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t.t if isinstance(t, X32) else t

    @staticmethod
    def _d(v):
        if isinstance(v, X32):
            return v.t.double()
        if torch.is_tensor(v):
            return v.double()
        return float(torch.tensor(float(v), dtype=torch.float32))

    @staticmethod
    def _f(v):
        if isinstance(v, X32):
            return v.t
        if torch.is_tensor(v):
            return v
        return float(torch.tensor(float(v), dtype=torch.float32))

    def __add__(self, o): return X32((X32._d(self) + X32._d(o)).float())
    def __radd__(self, o): return X32((X32._d(o) + X32._d(self)).float())
    def __sub__(self, o): return X32((X32._d(self) - X32._d(o)).float())
    def __rsub__(self, o): return X32((X32._d(o) - X32._d(self)).float())
    def __mul__(self, o): return X32((X32._d(self) * X32._d(o)).float())
    def __rmul__(self, o): return X32((X32._d(o) * X32._d(self)).float())
    def __truediv__(self, o): return X32((X32._d(self) / X32._d(o)).float())
    def __rtruediv__(self, o): return X32((X32._d(o) / X32._d(self)).float())
    def __neg__(self): return X32(-self.t)
    def __lt__(self, o): return self.t < X32._f(o)
    def __le__(self, o): return self.t <= X32._f(o)
    def __gt__(self, o): return self.t > X32._f(o)
    def __ge__(self, o): return self.t >= X32._f(o)
    def sqrt(self): return X32(torch.sqrt(self.t.double()).float())
    def expand(self, shape): return X32(self.t.expand(shape))
    def unsqueeze(self, d): return X32(self.t.unsqueeze(d))
    @property
    def shape(self): return self.t.shape


def xwhere(c, a, b):
    a, b = X32._f(a), X32._f(b)
    if not torch.is_tensor(a):
        a = torch.full_like(b, a)
    if not torch.is_tensor(b):
        b = torch.full_like(a, b)
    return X32(torch.where(c, a, b))


def xmax(a, s):
    """fmaxf(a, s) for a scalar s (NaN in ``a`` yields ``s``, like CUDA fmaxf)."""
    s = X32._f(s)
    return X32(torch.where(a.t >= s, a.t, torch.full_like(a.t, s)))


# ----------------------------------------------------------------------------
# A.0  pose preparation (reference models.py:207-211, 285-286; geometry.py:404)
# ----------------------------------------------------------------------------
def prepare_cameras(inp):
    """4x4 algebra the reference performs with torch.inverse/matmul.

    Returns a dict of fp32 tensors:
      Q     (b,n,4,4) inv(C) @ q            query cam2world in each ctx frame (models.py:208)
      Cself (b,n,4,4) inv(C) @ C            ~identity, passed to get_3d_point_epipolar (:207,283)
      Rel   (b,n,n,4,4) Rel[:,k] = inv(C[:,k]) @ C   ctx-j -> ctx-k (:285-286)
      qinv  (b,4,4)   inv(query cam2world)  (geometry.py:404 via models.py:586)
      K     (b,n,4,4) context intrinsics, Kq (b,4,4) query intrinsics
    """
    C = inp["context"]["cam2world"]
    q = inp["query"]["cam2world"]
    Cinv = torch.inverse(C)
    out = {
        "Q": torch.matmul(Cinv, q),
        "Cself": torch.matmul(Cinv, C),
        "Rel": torch.stack([torch.matmul(torch.inverse(C[:, k:k + 1]), C) for k in range(C.shape[1])], dim=1),
        "qinv": torch.inverse(q[:, 0]),
        "K": inp["context"]["intrinsics"].clone(),
        "Kq": inp["query"]["intrinsics"][:, 0].clone(),
    }
    return out


# ----------------------------------------------------------------------------
# helpers: fixed-order scalar arithmetic on tensors
# ----------------------------------------------------------------------------
def _dot3(a0, a1, a2, b0, b1, b2):
    """(a0*b0 + a1*b1) + a2*b2, each op rounded separately."""
    return (a0 * b0 + a1 * b1) + a2 * b2


def _norm3(x, y, z):
    q = (x * x + y * y) + z * z
    return q.sqrt() if isinstance(q, X32) else torch.sqrt(q)


def _cross(a, b):
    """a x b, components as (mul, mul, sub) — torch.cross (geometry.py:243)."""
    return (a[1] * b[2] - a[2] * b[1],
            a[2] * b[0] - a[0] * b[2],
            a[0] * b[1] - a[1] * b[0])


def _ray_through_pixel(u, v, fx, fy, cx, cy, M):
    """Unit direction (in the frame M maps into) of the ray through pixel
    (u, v), and its Plücker moment.  reference geometry.py:236-245 ->
    get_ray_directions :426-433 -> world_from_xy_depth :409-419 -> lift
    :353-371.

    M: tuple of 12 broadcastable tensors, rows 0..2 of the 4x4 cam2world.
    The einsum at geometry.py:417 is a K=4 dot; order used here:
    ((M0*x + M1*y) + M2*1) + M3*1.
    """
    u, v, fx, fy, cx, cy = (X32(t) for t in (u, v, fx, fy, cx, cy))
    M = tuple(X32(t) for t in M)
    xl = (u - cx) / fx            # * z with z == 1 is exact (geometry.py:365)
    yl = (v - cy) / fy
    p = []
    for r in range(3):
        m0, m1, m2, m3 = M[4 * r:4 * r + 4]
        p.append(((m0 * xl + m1 * yl) + m2) + m3)
    ox, oy, oz = M[3], M[7], M[11]
    dx, dy, dz = p[0] - ox, p[1] - oy, p[2] - oz
    nrm = xmax(_norm3(dx, dy, dz), 1e-12)                  # F.normalize eps (geometry.py:432)
    d = (dx / nrm, dy / nrm, dz / nrm)
    m = _cross((ox, oy, oz), d)
    return tuple(t.t for t in d), tuple(t.t for t in m)


def _rows(M):
    """(…,4,4) -> 12 tensors (rows 0..2), each with shape (…, 1) for ray broadcast."""
    return tuple(M[..., r, c].unsqueeze(-1) for r in range(3) for c in range(4))


# ----------------------------------------------------------------------------
# A.1  query ray in each context frame  (models.py:213-217)
# ----------------------------------------------------------------------------
def ray_setup(cams, uv):
    """uv (b,R,2) pixel (x,y).  Returns d, m (each tuple of 3 (b,n,R)) and o (b,n,3)."""
    Q = cams["Q"]                                           # (b,n,4,4)
    Kq = cams["Kq"]                                         # (b,4,4)
    fx = Kq[:, 0, 0][:, None, None]
    fy = Kq[:, 1, 1][:, None, None]
    cx = Kq[:, 0, 2][:, None, None]
    cy = Kq[:, 1, 2][:, None, None]
    u = uv[..., 0][:, None, :]                              # (b,1,R)
    v = uv[..., 1][:, None, :]
    d, m = _ray_through_pixel(u, v, fx, fy, cx, cy, _rows(Q))
    return d, m, Q[..., :3, 3]


# ----------------------------------------------------------------------------
# A.2  epipolar segment  (models.py:226-258 -> epipolar.py:175-253)
# ----------------------------------------------------------------------------
_EPS_LO = -1e-6           # epipolar.py:30,40  (python scalars are cast to fp32 by torch)
_EPS_HI = 1 + 1e-6


def _in_bounds(x, y):
    """epipolar.py:28-35 (NaN compares false)."""
    return (x >= _EPS_LO) & (y >= _EPS_LO) & (x <= _EPS_HI) & (y <= _EPS_HI)


def _project_norm(px, py, pz, Kn):
    """epipolar.py:23-26: p/(p.z+1e-8) then the 3x3 einsum, order (k0*x + k1*y) + k2*z."""
    den = pz + 1e-8
    qx, qy, qz = px / den, py / den, pz / den
    x = (Kn[0][0] * qx + Kn[0][1] * qy) + Kn[0][2] * qz
    y = (Kn[1][0] * qx + Kn[1][1] * qy) + Kn[1][2] * qz
    return x, y


def epipolar_segment(cams, d, o, H):
    """Clip the query ray to each context image.  Returns start, end (b,n,R,2)
    in grid coords [-1,1] after the NaN/Inf scrub, and overlaps (b,n,R) bool."""
    K = cams["K"]
    # intrinsics_norm: rows 0 AND 1 divided by H (models.py:228)
    Kn = [[(X32(K[..., r, c]) / float(H)).unsqueeze(-1) for c in range(3)] for r in range(2)]
    ox, oy, oz = (X32(o[..., i].unsqueeze(-1)) for i in range(3))       # (b,n,1)
    dx, dy, dz = (X32(t) for t in d)
    oxyz = (ox, oy, oz)
    dxyz = (dx, dy, dz)
    shape = dx.shape
    ts, xs, ys, vs = [], [], [], []
    for dim, val in ((0, 0.0), (0, 1.0), (1, 0.0), (1, 1.0)):           # epipolar.py:196-201
        od = 1 - dim
        fs, fo = Kn[dim][dim], Kn[od][od]
        cs, co = Kn[dim][2], Kn[od][2]
        os_, oo = oxyz[dim], oxyz[od]
        ds_, do = dxyz[dim], dxyz[od]
        c = (val - cs) / fs                                            # epipolar.py:99
        t = (c * oz - os_) / (ds_ - c * dz)                            # :103-105
        num = fo * (oo * (c * dz - ds_) + do * (os_ - c * oz))         # :109
        den = dz * os_ - ds_ * oz                                      # :110
        other = co + num / den                                         # :111
        same = X32(torch.full(shape, val, dtype=other.t.dtype, device=other.t.device))
        x, y = (same, other) if dim == 0 else (other, same)
        zz = oz + t * dz                                               # :116 (z component)
        valid = _in_bounds(x, y) & (zz > _EPS_LO)                      # :121
        ts.append(t.t.expand(shape)); xs.append(x.t.expand(shape)); ys.append(y.t.expand(shape)); vs.append(valid.expand(shape))

    def reduce(kind):                                                  # epipolar.py:125-149
        lowest = INF if kind == "min" else -INF
        bt = torch.where(vs[0], ts[0], torch.full_like(ts[0], lowest))
        bx, by, bv = xs[0], ys[0], vs[0]
        for i in range(1, 4):
            ti = torch.where(vs[i], ts[i], torch.full_like(ts[i], lowest))
            take = (ti < bt) if kind == "min" else (ti > bt)           # strict: first index wins ties
            bt = torch.where(take, ti, bt)
            bx = torch.where(take, xs[i], bx)
            by = torch.where(take, ys[i], by)
            bv = torch.where(take, vs[i], bv)
        return bx, by, bv
    fminx, fminy, fminv = reduce("min")
    fmaxx, fmaxy, fmaxv = reduce("max")

    # projection at t = 0 (epipolar.py:212-221)
    depth_zero = (oz < 1e-6).expand(shape)
    at_cam = (_norm3(ox, oy, oz) < 1e-6).expand(shape)
    px = xwhere(at_cam, dx, ox.expand(shape))
    py = xwhere(at_cam, dy, oy.expand(shape))
    pz = xwhere(at_cam, dz, oz.expand(shape))
    x0, y0 = _project_norm(px, py, pz, Kn)
    v0 = _in_bounds(x0, y0) & (pz > _EPS_LO)
    v0 = v0 & ~(depth_zero & ~at_cam)
    # projection at t = inf (epipolar.py:226-230)
    xi, yi = _project_norm(dx, dy, dz, Kn)
    vi = _in_bounds(xi, yi) & (dz > _EPS_LO)
    # merge (epipolar.py:241-251)
    minx = xwhere(v0, x0, fminx); miny = xwhere(v0, y0, fminy); minv = v0 | fminv
    maxx = xwhere(vi, xi, fmaxx); maxy = xwhere(vi, yi, fmaxy); maxv = vi | fmaxv
    overlaps = minv & maxv

    def to_grid(c):                                                    # models.py:246-252
        g = ((c - 0.5) * 2).t
        return torch.where(torch.isfinite(g), g, torch.zeros_like(g))
    start = torch.stack([to_grid(minx), to_grid(miny)], dim=-1)
    end = torch.stack([to_grid(maxx), to_grid(maxy)], dim=-1)
    return start, end, overlaps


# ----------------------------------------------------------------------------
# A.3  line samples (models.py:261, 271-275)
# ----------------------------------------------------------------------------
def line_samples(start, end, interval):
    diff = X32(end[..., None, :]) - X32(start[..., None, :])
    return (X32(start[..., None, :]) + diff * X32(interval[None, None, None, :, None])).t   # (b,n,R,P,2)


# ----------------------------------------------------------------------------
# A.4 / A.6  bilinear gathers (PyTorch CUDA grid_sample formulas,
#            ATen/native/cuda/GridSampler.cuh:23-31,56-59,139-168)
# ----------------------------------------------------------------------------
def bilinear_taps(gx, gy, w, h, border):
    """Returns ix_nw, iy_nw (int64) and the 4 weights (nw, ne, sw, se)."""
    ix = (((X32(gx) + 1.0) * float(w) - 1.0) / 2.0).t
    iy = (((X32(gy) + 1.0) * float(h) - 1.0) / 2.0).t
    if border:
        ix = torch.clamp(torch.where(torch.isnan(ix), torch.zeros_like(ix), ix), 0, w - 1)
        iy = torch.clamp(torch.where(torch.isnan(iy), torch.zeros_like(iy), iy), 0, h - 1)
    bad = lambda c: (c > 2147483646.0) | (c < -2147483648.0) | ~torch.isfinite(c)
    ix = torch.where(bad(ix), torch.full_like(ix, -100.0), ix)
    iy = torch.where(bad(iy), torch.full_like(iy, -100.0), iy)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    wx1 = ix - x0
    wx0 = (x0 + 1) - ix
    wy1 = iy - y0
    wy0 = (y0 + 1) - iy
    return x0.long(), y0.long(), (wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1)


def gather_bilinear(maps, gx, gy, border):
    """maps: list of (B,C,h,w); gx, gy (B,R,P).  Returns (B,R,P,sum C)."""
    outs = []
    B = gx.shape[0]
    for zmap in maps:
        _, C, h, w = zmap.shape
        x0, y0, wts = bilinear_taps(gx, gy, w, h, border)
        nhwc = zmap.permute(0, 2, 3, 1).reshape(B, h * w, C)
        acc = torch.zeros(B, gx.shape[1], gx.shape[2], C, dtype=zmap.dtype, device=zmap.device)
        for (ddx, ddy), wt in zip(((0, 0), (1, 0), (0, 1), (1, 1)), wts):
            xx, yy = x0 + ddx, y0 + ddy
            inb = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).reshape(B, -1)
            vals = torch.gather(nhwc, 1, idx[..., None].expand(-1, -1, C)).reshape(*gx.shape, C)
            acc = acc + vals * (wt * inb.to(wt.dtype))[..., None]
        outs.append(acc)
    return torch.cat(outs, dim=-1)


# ----------------------------------------------------------------------------
# A.5  triangulated point per sample, fp64 (geometry.py:98-162)
# ----------------------------------------------------------------------------
def triangulate(cams, d, m, pixel_val, H, W):
    """pt (b,n,R,P,3) fp32: closest point on the query ray to the ray through
    the context pixel, in that context's frame.  Also returns px, py."""
    K = cams["K"]
    fx = K[..., 0, 0][..., None, None]; fy = K[..., 1, 1][..., None, None]
    cx = K[..., 0, 2][..., None, None]; cy = K[..., 1, 2][..., None, None]
    px = ((X32(pixel_val[..., 0]) + 1.0) / 2.0 * float(W - 1)).t      # geometry.py:101
    py = ((X32(pixel_val[..., 1]) + 1.0) / 2.0 * float(H - 1)).t      # :100
    M = tuple(t.unsqueeze(-1) for t in _rows(cams["Cself"]))          # (b,n,1,1)
    l2, m2 = _ray_through_pixel(px, py, fx, fy, cx, cy, M)            # :108
    D = torch.float64
    l1 = [t.unsqueeze(-1).to(D) for t in d]                           # (b,n,R,1)
    m1 = [t.unsqueeze(-1).to(D) for t in m]
    l2 = [t.to(D) for t in l2]
    m2 = [t.to(D) for t in m2]
    n = _cross(l1, l2)                                                # :142
    a = _cross(l2, n)                                                 # :143
    t1 = _cross(m1, a)                                                # :145 (negated below)
    s = _dot3(m2[0], m2[1], m2[2], n[0], n[1], n[2])                  # :147
    nn = _norm3(n[0], n[1], n[2])
    cd = nn * nn + 1e-12                                              # :149
    pt = []
    for i in range(3):
        p = (-t1[i] + s * l1[i]) / cd                                 # :151
        p = torch.where(torch.isfinite(p), p, torch.zeros_like(p))    # :126-127
        pt.append(p.to(torch.float32))
    return torch.stack(pt, dim=-1), px, py


# ----------------------------------------------------------------------------
# A.6  cross-view reprojection (models.py:285-325; geometry.py:374-393; utils/util.py:16-19)
# ----------------------------------------------------------------------------
def _xform_point(T, pt):
    """encode_relative_point (models.py:30-39): broadcast-multiply-sum with K=4,
    order ((T0*x + T1*y) + T2*z) + T3."""
    x, y, z = pt[..., 0], pt[..., 1], pt[..., 2]
    out = []
    for r in range(3):
        t = [T[..., r, c][..., None, None] for c in range(4)]
        out.append(((t[0] * x + t[1] * y) + t[2] * z) + t[3])
    return torch.stack(out, dim=-1)


def reproject(cams, pt, H, W):
    """Returns pt_v0, pt_v1 (b,n,R,P,3): pt expressed in view-0 / view-1 frames
    (raw, before the NaN scrub), and cross grid coords gxc, gyc (b,n,R,P):
    for ctx-0 rows the projection into view 1, for ctx-1 rows into view 0."""
    Rel = cams["Rel"]                                                 # (b,2,n,4,4)
    K = cams["K"]
    pt_v0 = _xform_point(Rel[:, 0], pt)
    pt_v1 = _xform_point(Rel[:, 1], pt)
    other = torch.stack([pt_v1[:, 0], pt_v0[:, 1]], dim=1)           # ctx0->view1, ctx1->view0
    Ko = torch.stack([K[:, 1], K[:, 0]], dim=1)                      # intrinsics of the *other* view
    fx = Ko[..., 0, 0][..., None, None]; fy = Ko[..., 1, 1][..., None, None]
    cx = Ko[..., 0, 2][..., None, None]; cy = Ko[..., 1, 2][..., None, None]
    X, Y, Z = other[..., 0], other[..., 1], other[..., 2]
    xp = fx * X / (Z + 1e-12) + cx                                    # geometry.py:386
    yp = fy * Y / (Z + 1e-12) + cy
    big = torch.full_like(xp, 1e10)
    xp = torch.where(torch.isfinite(xp), xp, big)                     # :390-391
    yp = torch.where(torch.isfinite(yp), yp, big)
    gxc = (xp / (W - 1)) * 2 - 1                                      # utils/util.py:17
    gyc = (yp / (H - 1)) * 2 - 1
    return pt_v0, pt_v1, gxc, gyc


# ----------------------------------------------------------------------------
# A.7 - A.12  per-sample MLPs, attention, colour MLP
# ----------------------------------------------------------------------------
def _lin(sd, name, x):
    w = sd[name + ".weight"]
    return F.linear(x, w.reshape(w.shape[0], -1), sd[name + ".bias"])


def local_coords(cams, d, o, pt, px, py):
    """16-channel geometric query feature (models.py:494-528; geometry.py:313-324)."""
    K = cams["K"]
    fx = K[..., 0, 0][..., None, None]; fy = K[..., 1, 1][..., None, None]
    cx = K[..., 0, 2][..., None, None]; cy = K[..., 1, 2][..., None, None]
    rx = (px - cx) / fx
    ry = (py - cy) / fy
    rz = torch.ones_like(rx)
    nrm = torch.clamp_min(_norm3(rx, ry, rz), 1e-12)
    cam = torch.stack([rx / nrm, ry / nrm, rz / nrm], dim=-1)
    oo = o[:, :, None, None, :]
    df = pt - oo
    depth = _norm3(df[..., 0], df[..., 1], df[..., 2])
    depth = torch.where(torch.isfinite(depth), depth, torch.full_like(depth, 1000000.0))
    dd = torch.stack([t.unsqueeze(-1).expand_as(px) for t in d], dim=-1)
    de = torch.stack([torch.tanh(depth), torch.tanh(depth / 10.), torch.tanh(depth / 100.),
                      torch.tanh(depth / 1000.)], dim=-1)
    return torch.cat([cam, torch.zeros_like(cam), dd, de, oo.expand_as(cam)], dim=-1)


def _nan_to_num(x):
    return torch.nan_to_num(x, 0)                                     # models.py:322-325


def render(sd, inp, z, H, W, P, interval=None, cams=None, no_sample=False, no_latent_concat=False):
    """Full hot path.  sd: renderer state_dict (fp32), inp: reference-style input
    dict, z: [z1,z2,z3] NCHW.  Returns dict with every out_dict entry of the
    reference (models.py:217-218,570-571,592-597,617-624) and the intermediates.

    Ablation branches (ORACLE ONLY in this round; the CUDA path runs the default flags):
      no_sample         epipolar line = the query ray's points at depths linspace(0.1, 10, P) projected
                        into each context (geometry.get_epipolar_lines_volumetric :165-187), no clipping
      no_latent_concat  no cross-view gather / per-sample encoder: the raw 576 gathered channels feed
                        latent_value / key_map directly (models.py:476-477,121-124)"""
    dev = z[0].device
    b, n = inp["context"]["cam2world"].shape[:2]
    assert n == 2
    uv = inp["query"]["uv"][:, 0]                                     # (b,R,2)
    R = uv.shape[1]
    if cams is None:
        cams = prepare_cameras(inp)
    if interval is None:
        interval = torch.linspace(0, 1, P, device=dev)                # models.py:261
    I = {}
    d, m, o = ray_setup(cams, uv)
    if no_sample:
        # geometry.py:165-187: same elementwise torch ops in the same order as the reference (plain fp32)
        ivl = torch.linspace(0.1, 10., P, device=dev)
        dvec = torch.stack(d, dim=-1)                                 # (b,n,R,3)
        pts = o[:, :, None, None, :] + ivl[None, None, None, :, None] * dvec[..., None, :]
        Kc = cams["K"]
        fx = Kc[..., 0, 0][..., None, None]; fy = Kc[..., 1, 1][..., None, None]
        cx = Kc[..., 0, 2][..., None, None]; cy = Kc[..., 1, 2][..., None, None]
        xp = fx * pts[..., 0] / (pts[..., 2] + 1e-12) + cx            # geometry.project :386-387
        yp = fy * pts[..., 1] / (pts[..., 2] + 1e-12) + cy
        big = torch.full_like(xp, 1e10)
        xp = torch.where(torch.isfinite(xp), xp, big)                 # :390-391
        yp = torch.where(torch.isfinite(yp), yp, big)
        pv = torch.stack([(xp / (W - 1)) * 2 - 1, (yp / (H - 1)) * 2 - 1], dim=-1)   # utils/util.py:16-19
        start, end = pv[..., 0, :], pv[..., -1, :]
        overlaps = ((pv < 1) & (pv > -1)).all(dim=-1).any(dim=-1)     # :185 ("no_intersect")
    else:
        start, end, overlaps = epipolar_segment(cams, d, o, H)
        pv = line_samples(start, end, interval)                       # (b,n,R,P,2)
    I["pixel_val"] = pv
    gx, gy = pv[..., 0], pv[..., 1]
    zmaps = [t.reshape(b, n, *t.shape[1:]) for t in z]
    flat = lambda t: t.reshape(b * n, *t.shape[2:])
    f_own = gather_bilinear(z, flat(gx), flat(gy), border=True).reshape(b, n, R, P, -1)
    pt, px, py = triangulate(cams, d, m, pv, H, W)
    if no_latent_concat:
        V = _lin(sd, "latent_value", f_own)                           # models.py:476-477,487
        loc = local_coords(cams, d, o, pt, px, py)
        out, I2 = _attention_and_colour(sd, cams, b, n, R, P, V, f_own, loc, pt, d, m, o, overlaps)
        I.update(I2)
        I.update(feat_primary=f_own, value=V, local=loc, pt=pt)
        out.update({"pixel_val": pv.reshape(b * n, R, P, 2), "uv": inp["query"]["uv"], "z": z,
                    "_overlaps": overlaps, "_start": start, "_end": end, "_cams": cams, "_I": I})
        return out
    pt_v0, pt_v1, gxc, gyc = reproject(cams, pt, H, W)
    I["grid_cross"] = torch.stack([gxc, gyc], dim=-1)
    # features of the OTHER view at the reprojected point (models.py:316-320)
    z_other = [torch.stack([t[:, 1], t[:, 0]], dim=1).reshape(b * n, *t.shape[2:]) for t in zmaps]
    f_oth = gather_bilinear(z_other, flat(gxc), flat(gyc), border=False).reshape(b, n, R, P, -1)
    I["feat_primary"], I["feat_cross"] = f_own, f_oth
    # per-row view-0 / view-1 features and points (models.py:330-342)
    ctx0 = torch.zeros(1, n, 1, 1, 1, dtype=torch.bool, device=dev); ctx0[:, 0] = True
    f_v0 = torch.where(ctx0, f_own, f_oth)
    f_v1 = torch.where(ctx0, f_oth, f_own)
    x_v0 = torch.cat([f_v0, torch.tanh(_nan_to_num(pt_v0) / 5.)], dim=-1)
    x_v1 = torch.cat([f_v1, torch.tanh(_nan_to_num(pt_v1) / 5.)], dim=-1)
    enc = lambda x: _lin(sd, "query_encode_latent_2", F.relu(_lin(sd, "query_encode_latent", x)))
    e0, e1 = enc(x_v0), enc(x_v1)
    interp = torch.cat([e0, e1], dim=-1)                              # (b,n,R,P,576)
    I["enc_v0"], I["enc_v1"] = e0, e1
    V = _lin(sd, "latent_value", interp)                              # models.py:487
    loc = local_coords(cams, d, o, pt, px, py)                        # :494-528
    out, I2 = _attention_and_colour(sd, cams, b, n, R, P, V, interp, loc, pt, d, m, o, overlaps)
    I.update(I2)
    I.update(value=V, local=loc, pt=pt)
    out.update({
        "pixel_val": pv.reshape(b * n, R, P, 2),
        "uv": inp["query"]["uv"],
        "z": z,
        "_overlaps": overlaps,
        "_start": start, "_end": end,
        "_cams": cams,
        "_I": I,
    })
    return out


def _attention_and_colour(sd, cams, b, n, R, P, V, interp_for_key, loc, pt, d, m, o, overlaps):
    """Shared tail of every n_view branch (models.py:487-621): K, geometric Q, two joint-softmax
    rounds over the ray's n*P samples, expected depth, colour MLP, white fill.  ``V`` already holds
    latent_value(interp); returns (out entries, intermediates)."""
    Kk = _lin(sd, "key_map_2", F.relu(_lin(sd, "key_map", interp_for_key)))   # :491
    Q1 = _lin(sd, "query_embed_2", F.relu(_lin(sd, "query_embed", loc)))      # :529

    def joint_softmax(s_):                                             # models.py:533-535
        sj = s_.permute(0, 2, 1, 3).reshape(b, R, n * P)
        a = F.softmax(sj, dim=-1)
        return a.reshape(b, R, n, P).permute(0, 2, 1, 3)
    s1 = (Kk * Q1).sum(-1) / 16.                                       # :532
    a1 = joint_softmax(s1)
    zsum = (V * a1[..., None]).sum(dim=3).sum(dim=1)                   # (b,R,L)  :537-540
    g = _lin(sd, "encode_latent", zsum)                                # :548
    qin = torch.cat([g[:, None, :, None, :].expand(-1, n, -1, P, -1), loc], dim=-1)   # :552
    Q2 = _lin(sd, "query_repeat_embed_2", F.relu(_lin(sd, "query_repeat_embed", qin)))
    s2 = (Q2 * Q1).sum(-1) / 16.                                       # :555
    a2 = joint_softmax(s2)
    zloc2 = (V * a2[..., None]).sum(dim=3) + zsum[:, None]             # :561  (b,n,R,L)
    zfin = zloc2.sum(dim=1)                                            # :564: every ctx row holds this sum
    at_max = a1.argmax(dim=-1)
    w3d = (a1[..., None] * torch.clamp(pt, -100, 100)).sum(dim=3).sum(dim=1)
    qi = cams["qinv"]
    zc = ((qi[:, 2, 0, None] * w3d[..., 0] + qi[:, 2, 1, None] * w3d[..., 1])
          + qi[:, 2, 2, None] * w3d[..., 2]) + qi[:, 2, 3, None]
    depth_ray = torch.clamp(zc, 0, 10)
    dd = torch.stack(d, dim=-1); mm = torch.stack(m, dim=-1)
    coords9 = torch.cat([dd, mm, o[:, :, None, :].expand(-1, -1, R, -1)], dim=-1)    # (b,n,R,9)
    cflat = coords9.permute(0, 2, 1, 3).reshape(b, R, n * 9)          # :602
    zcat = torch.cat([zfin] * n, dim=-1)                               # :605-606 (identical copies per ctx)
    x = _lin(sd, "phi.lin_in", cflat)
    for i in range(3):
        x = x + _lin(sd, f"phi.lin_z.{i}", zcat)
        net = _lin(sd, f"phi.blocks.{i}.fc_0", F.relu(x))
        x = x + _lin(sd, f"phi.blocks.{i}.fc_1", F.relu(net))
    rgb = _lin(sd, "phi.lin_out", F.relu(x))
    valid = overlaps.any(dim=1).float()
    rgb = rgb * valid[..., None] + 1 * (1 - valid[..., None])
    out = {"rgb": rgb.reshape(b, 1, R, 3), "valid_mask": valid[..., None], "depth_ray": depth_ray[..., None],
           "at_wt": a1.reshape(b * n, R, P), "at_wts": [a1.reshape(b * n, R, P)],
           "at_wt_max": at_max.reshape(b * n, R, 1), "coords": coords9.reshape(b * n, R, 9)}
    I = dict(key=Kk, q1=Q1, s1=s1, at_wt=a1, zsum=zsum, g=g, q2=Q2, s2=s2, at_wt2=a2, z_final=zfin)
    return out, I


def render_single_view(sd, inp, z, H, W, P, interval=None, cams=None):
    """``n_view = 1`` branch of the reference forward (models.py:478-485 + the shared tail): one
    context view, no cross-view gather; the 576 gathered channels and
    [tanh(pt/5), tanh(pt/100)] go through ``update_val_merge`` (582 -> 576, no ReLU), V / K are
    576-wide, the softmax runs over the P samples of the single line, phi sees 9 + 576 inputs.
    ORACLE ONLY in this round: the CUDA path covers n_view = 2."""
    dev = z[0].device
    b, n = inp["context"]["cam2world"].shape[:2]
    assert n == 1
    uv = inp["query"]["uv"][:, 0]
    R = uv.shape[1]
    if cams is None:
        cams = prepare_cameras(inp)
    if interval is None:
        interval = torch.linspace(0, 1, P, device=dev)
    d, m, o = ray_setup(cams, uv)
    start, end, overlaps = epipolar_segment(cams, d, o, H)
    pv = line_samples(start, end, interval)                           # (b,1,R,P,2)
    gx, gy = pv[..., 0], pv[..., 1]
    flat = lambda t: t.reshape(b * n, *t.shape[2:])
    f_own = gather_bilinear(z, flat(gx), flat(gy), border=True).reshape(b, n, R, P, -1)
    pt, px, py = triangulate(cams, d, m, pv, H, W)                    # NaN/Inf already scrubbed (:126-127,481)
    pt_context = torch.cat([torch.tanh(pt / 5.), torch.tanh(pt / 100.)], dim=-1)      # :483
    interp = _lin(sd, "update_val_merge", torch.cat([f_own, pt_context], dim=-1))     # :484-485
    V = _lin(sd, "latent_value", interp)                              # :487
    loc = local_coords(cams, d, o, pt, px, py)
    out, I = _attention_and_colour(sd, cams, b, n, R, P, V, interp, loc, pt, d, m, o, overlaps)
    I.update(pixel_val=pv, feat_primary=f_own, merged=interp, value=V, local=loc, pt=pt)
    out.update({"pixel_val": pv.reshape(b * n, R, P, 2), "uv": inp["query"]["uv"], "z": z,
                "_overlaps": overlaps, "_cams": cams, "_I": I})
    return out


def _project_to_grid(ptk, Kj, H, W):
    """geometry.project (geometry.py:374-393) with view j's intrinsics, then
    util.normalize_for_grid_sample (utils/util.py:16-19).  ptk (b,R,P,3), Kj (b,4,4)."""
    fx = Kj[:, 0, 0][:, None, None]; fy = Kj[:, 1, 1][:, None, None]
    cx = Kj[:, 0, 2][:, None, None]; cy = Kj[:, 1, 2][:, None, None]
    X, Y, Z = ptk[..., 0], ptk[..., 1], ptk[..., 2]
    xp = fx * X / (Z + 1e-12) + cx
    yp = fy * Y / (Z + 1e-12) + cy
    big = torch.full_like(xp, 1e10)
    xp = torch.where(torch.isfinite(xp), xp, big)
    yp = torch.where(torch.isfinite(yp), yp, big)
    return (xp / (W - 1)) * 2 - 1, (yp / (H - 1)) * 2 - 1


def render_three_views(sd, inp, z, H, W, P, interval=None, cams=None):
    """``n_view = 3`` branch of the reference forward (models.py:345-475 + the shared tail).
    What the code does (followed literally, including its frame bookkeeping): with
    ptv[a][j] = the samples of context j's epipolar line expressed in view a's frame
    (models.py:354-382), the row (ray r, sample p) of context a is encoded from
      k = 0: its own gathered features and tanh(ptv[a][a] / 5)                       (:436-439)
      k = 1, 2 (the other contexts j in ascending order): view j's maps sampled (zero padding)
             at project(ptv[a][j], K_j) - i.e. the (r, p) sample of context j's OWN line - and
             tanh(ptv[a][j] / 5)                                                      (:385-423, 437)
    each through query_encode_latent(_2); the three 288-vectors are interleaved channel-major
    (channel c of part k at index 3c + k: ``cat(dim=2).flatten(1, 2)``, :444-446).
    ORACLE ONLY in this round: the CUDA path covers n_view = 2."""
    dev = z[0].device
    b, n = inp["context"]["cam2world"].shape[:2]
    assert n == 3
    uv = inp["query"]["uv"][:, 0]
    R = uv.shape[1]
    if cams is None:
        cams = prepare_cameras(inp)
    if interval is None:
        interval = torch.linspace(0, 1, P, device=dev)
    d, m, o = ray_setup(cams, uv)
    start, end, overlaps = epipolar_segment(cams, d, o, H)
    pv = line_samples(start, end, interval)                           # (b,3,R,P,2)
    gx, gy = pv[..., 0], pv[..., 1]
    flat = lambda t: t.reshape(b * n, *t.shape[2:])
    f_own = gather_bilinear(z, flat(gx), flat(gy), border=True).reshape(b, n, R, P, -1)
    pt, px, py = triangulate(cams, d, m, pv, H, W)                    # (b,3,R,P,3), ctx j's samples in frame j
    Rel, K = cams["Rel"], cams["K"]                                   # Rel[:, a, j]: frame j -> frame a
    zmaps = [t.reshape(b, n, *t.shape[1:]) for t in z]
    enc = lambda x: _lin(sd, "query_encode_latent_2", F.relu(_lin(sd, "query_encode_latent", x)))
    rows = []
    for a in range(n):
        ptv_a = _xform_point(Rel[:, a], pt)                           # (b,3,R,P,3): [j] = ctx j's samples in frame a
        parts = [enc(torch.cat([f_own[:, a], torch.tanh(_nan_to_num(ptv_a[:, a]) / 5.)], dim=-1))]
        for j in range(n):
            if j == a:
                continue
            gxc, gyc = _project_to_grid(ptv_a[:, j], K[:, j], H, W)  # :394-401 (intrinsics of view j)
            f_j = gather_bilinear([t[:, j] for t in zmaps], gxc, gyc, border=False)        # :403-404
            parts.append(enc(torch.cat([f_j, torch.tanh(_nan_to_num(ptv_a[:, j]) / 5.)], dim=-1)))
        rows.append(torch.stack(parts, dim=-1).flatten(-2, -1))       # (b,R,P,288*3), index 3c + k
    interp = torch.stack(rows, dim=1)                                 # (b,3,R,P,864)  :473
    V = _lin(sd, "latent_value", interp)
    loc = local_coords(cams, d, o, pt, px, py)
    out, I = _attention_and_colour(sd, cams, b, n, R, P, V, interp, loc, pt, d, m, o, overlaps)
    I.update(pixel_val=pv, feat_primary=f_own, interp=interp, value=V, local=loc, pt=pt)
    out.update({"pixel_val": pv.reshape(b * n, R, P, 2), "uv": inp["query"]["uv"], "z": z,
                "_overlaps": overlaps, "_cams": cams, "_I": I})
    return out


def primary_taps(pixel_val, w, h):
    """Integer NW taps of the primary (border) gather for a map of size (h,w)."""
    x0, y0, _ = bilinear_taps(pixel_val[..., 0], pixel_val[..., 1], w, h, border=True)
    return x0, y0


def psnr(a, b):
    """eval metric of the reference (experiment_scripts/eval_realestate10k.py:74-75,181)."""
    a = (a + 1) / 2
    b = (b + 1) / 2
    return -10.0 * math.log10(max(float(((a - b) ** 2).mean()), 1e-20))


def render_grad(sd, inp, z, H, W, P, g_rgb=None, g_depth=None, cams=None):
    """Gradients of  L = sum(rgb * g_rgb) + sum(depth_ray * g_depth)  w.r.t. every tensor of
    ``sd`` and the three feature maps, by torch autograd through ``render`` — the restatement
    of the reference's ``train_loss.backward()`` (training.py:125) for the renderer.  The
    geometry stages run on detached exactly-rounded fp32 (no parameter lies upstream of the
    sample coordinates; the reference detaches pt/depth, models.py:327-328,516).
    Returns (out_dict, {name: grad}, [dz1, dz2, dz3])."""
    sd_r = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    z_r = [t.detach().clone().requires_grad_(True) for t in z]
    with torch.enable_grad():
        out = render(sd_r, inp, z_r, H, W, P, cams=cams)
        loss = 0.0
        if g_rgb is not None:
            loss = loss + (out["rgb"] * g_rgb).sum()
        if g_depth is not None:
            loss = loss + (out["depth_ray"] * g_depth).sum()
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd_r.items()}
    return out, grads, [t.grad for t in z_r]
