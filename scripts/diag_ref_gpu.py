"""Diagnostic: who disagrees with whom at 256x256 / 64 samples?  (this repo's CUDA path, the unmodified
reference on the GPU with TF32 off, the unmodified reference on the host, the oracle port)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer
from oracle import car_oracle as orc, ref_loader

DEV = "cuda:0"
b, H, P, seed, rays = 1, 256, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 54, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
inp = synthetic.make_inputs(b, H, H, seed=seed, rays=rays)
z = synthetic.make_features(b, H, seed=seed)
sd = synthetic.make_state_dict(seed=seed)
inp_d = synthetic.to_device(inp, DEV)
z_d = [t.to(DEV) for t in z]
with ref_loader.strict_fp32():
    m_gpu = ref_loader.build_model(sd, H, P, device=DEV)
    ref_gpu = ref_loader.render(m_gpu, inp_d, z_d, chunk_rays=2048)
with ref_loader.host_mode():
    m_cpu = ref_loader.build_model(sd, H, P, device="cpu")
    ref_cpu = ref_loader.render(m_cpu, inp, z, chunk_rays=2048)
with torch.no_grad():
    port = orc.render(sd, inp, z, H, H, P)
outs = {}
for prec in ("fp32", "fp32_simt"):
    m = CrossAttentionRenderer(n_view=2, npoints=P, precision=prec).to(DEV).eval()
    m.load_state_dict(sd, strict=False); m.H = m.W = H
    with torch.no_grad():
        outs[prec] = m(inp_d, z=z_d)
        # same kernels fed with the CPU-prepared 4x4s
        cams = {k: v.to(DEV).contiguous() for k, v in orc.prepare_cameras(inp).items()}
        outs[prec + "/cpu-cams"] = m.render_prepared(cams, inp["query"]["uv"][:, 0].contiguous().to(DEV),
                                                     torch.linspace(0, 1, P, device=DEV), z_d, b, rays)
torch.cuda.synchronize()
runs = {"ref_gpu": ref_gpu, "ref_cpu": ref_cpu, "port": port, **outs}
rgb = {k: v["rgb"].detach().cpu().reshape(-1, 3) for k, v in runs.items()}
scale = float(rgb["ref_gpu"].abs().max())
names = list(rgb)
print(f"seed {seed}, {rays} rays; rgb scale {scale:.3f}; max |a-b| / scale, (rays above 1e-4)")
for i, a in enumerate(names):
    for bname in names[i + 1:]:
        d = (rgb[a] - rgb[bname]).abs().amax(dim=-1) / scale
        print(f"  {a:18s} vs {bname:18s}: {float(d.max()):.3e}  ({int((d > 1e-4).sum())})  median {float(d.median()):.2e}")
d = (rgb["fp32"] - rgb["ref_gpu"]).abs().amax(dim=-1) / scale
worst = d.topk(5).indices
pv = {k: v["pixel_val"].detach().cpu() for k, v in runs.items()}
aw = {k: v["at_wt"].detach().cpu() for k, v in runs.items()}
for r in worst.tolist():
    print(f"ray {r}: err(fp32 vs ref_gpu) {float(d[r]):.2e}; ref_gpu vs ref_cpu {float((rgb['ref_gpu'][r] - rgb['ref_cpu'][r]).abs().max() / scale):.2e}; "
          f"fp32 vs port {float((rgb['fp32'][r] - rgb['port'][r]).abs().max() / scale):.2e}; "
          f"pixel_val diff (fp32 vs ref_gpu) {float((pv['fp32'][:, r] - pv['ref_gpu'][:, r]).abs().max()):.2e}; "
          f"at_wt max {float(aw['ref_gpu'][:, r].max()):.3f} diff {float((aw['fp32'][:, r] - aw['ref_gpu'][:, r]).abs().max()):.2e}")
# conditioning: near the epipole the triangulated point (geometry.py:132-162) runs off to +-infinity and
# flips sign; is every ray above 1e-4 one whose samples include such a point?
pt = port["_I"]["pt"]                                     # (b,n,R,P,3)
ptmax = pt.abs().amax(dim=(1, 3, 4)).reshape(-1)         # per ray
for name_a, name_b in (("fp32", "ref_gpu"), ("ref_gpu", "ref_cpu"), ("fp32", "port")):
    dd = (rgb[name_a] - rgb[name_b]).abs().amax(dim=-1) / scale
    bad = dd > 1e-4
    print(f"{name_a} vs {name_b}: {int(bad.sum())} rays above 1e-4; their max|pt|: {sorted(ptmax[bad].tolist())[:20]}")
for thr in (1e2, 1e3, 1e4, 1e5):
    print(f"rays with max|pt| > {thr:g}: {int((ptmax > thr).sum())} of {ptmax.numel()}")
