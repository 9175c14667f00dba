"""Diagnostic: where do CUDA sample coordinates differ from the fixed-order oracle?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer
from oracle import car_oracle as orc
from golden_util import ulp_diff

b, H, Ht, P = 2, 64, 64, 32
inp = synthetic.make_inputs(b, H, Ht, seed=7, mode="default")
z = synthetic.make_features(b, H, seed=7)
sd = synthetic.make_state_dict(seed=7)
cams = orc.prepare_cameras(inp)
interval = torch.linspace(0, 1, P)
uv = inp["query"]["uv"][:, 0]

def oracle_on(dev):
    c = {k: v.to(dev) for k, v in cams.items()}
    d, m, o = orc.ray_setup(c, uv.to(dev))
    s, e, ov = orc.epipolar_segment(c, d, o, H)
    pv = orc.line_samples(s, e, interval.to(dev))
    return [t.cpu() for t in (torch.stack(d, -1), torch.stack(m, -1), s, e, pv)]

cpu = oracle_on("cpu")
gpu = oracle_on("cuda")
names = ["d", "m", "start", "end", "pixel_val"]
print("torch-CPU oracle vs torch-CUDA oracle (same code):")
for n, a, g in zip(names, cpu, gpu):
    u = ulp_diff(a, g)
    print(f"  {n:10s} mismatches {int((u>0).sum()):8d} / {u.numel()}  max ulp {int(u.max())}")
m = CrossAttentionRenderer(n_view=2, npoints=P, precision="fp32_simt").cuda()
m.load_state_dict(sd, strict=False); m.H = m.W = H
camsd = {k: v.cuda().contiguous() for k, v in cams.items()}
out = m.render_prepared(camsd, uv.contiguous().cuda(), interval.cuda(), [t.cuda() for t in z], b, uv.shape[1])
co = out["coords"].cpu().reshape(b, 2, -1, 9)
pv = out["pixel_val"].reshape(b, 2, -1, P, 2)
for tag, ref in (("CPU", cpu), ("CUDA", gpu)):
    print(f"kernel vs torch-{tag} oracle:")
    for n, a, g in (("d", ref[0], co[..., 0:3]), ("m", ref[1], co[..., 3:6]),
                    ("start", ref[2], pv[..., 0, :]), ("end", ref[3], pv[..., -1, :]), ("pixel_val", ref[4], pv)):
        u = ulp_diff(a, g)
        print(f"  {n:10s} mismatches {int((u>0).sum()):8d} / {u.numel()}  max ulp {int(u.max())}")
# which coordinate components / interior only?
u = ulp_diff(cpu[4], pv)
bad = (u > 0)
print("bad by sample index k:", bad.sum(dim=(0, 1, 2, 4)).tolist())
i = bad.nonzero()[:5]
for r in i:
    bb, j, rr, k, c = r.tolist()
    print(r.tolist(), float(cpu[4][bb, j, rr, k, c]), float(pv[bb, j, rr, k, c]), "start", float(cpu[2][bb, j, rr, c]), "end", float(cpu[3][bb, j, rr, c]), "iv", float(interval[k]))
