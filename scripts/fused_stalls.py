"""Where does each role of the fused kernel wait?  (pair 0, leader CTA, cycles)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic, _lib
from cross_attention_renderer_b200.models import CrossAttentionRenderer
lib = _lib.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
b, H, P = 1, 256, 64
inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=1), "cuda")
z = [t.cuda() for t in synthetic.make_features(b, H, seed=1)]
m = CrossAttentionRenderer(n_view=2, npoints=P, precision=prec).cuda()
m.load_state_dict(synthetic.make_state_dict(0), strict=False); m.H = m.W = H; m.pixel_val_to_cpu = False
with torch.no_grad():
    m(inp, z=z); torch.cuda.synchronize()
    stats = torch.zeros(64, dtype=torch.int64, device="cuda")
    lib.car_debug_set_fused_stats(stats.data_ptr())
    m(inp, z=z); torch.cuda.synchronize()
    lib.car_debug_set_fused_stats(None)
s = stats.cpu().tolist()
names = {0: "mma:a1_empty", 1: "mma:x_full", 2: "mma:b_full(g1)", 3: "mma:a3_empty", 4: "mma:h_full", 5: "mma:b_full(g3)", 6: "mma:TOTAL",
         8: "epi:a1_full", 9: "epi:h_empty", 10: "epi:a3_full", 11: "epi:compute+sts", 12: "epi:fence", 13: "epi:arrive", 14: "epi:a3 drain", 15: "epi:TOTAL",
         16: "prod:x_empty", 17: "prod:TOTAL", 18: "prod:bar", 19: "prod:fence+arrive", 20: "tma:b_empty", 21: "tma:TOTAL", 22: "epi:a3 flush (bar+copy)"}
rays_pair0 = (b * H * H + 73) // 74
print(f"precision {prec}: pair-0 rays ~{rays_pair0}")
for k, n in names.items():
    tot = s[6] if k < 8 else s[15] if (k < 16 or k == 22) else s[17] if k < 20 else s[21]
    print(f"  {n:18s} {100.0 * s[k] / max(1, tot):6.1f}%   {s[k] / rays_pair0:10.0f} cyc/ray")

tn = ["drain(hidden)", "wait MMA (scores)", "scores", "softmax + outputs", "post weights (a_empty)", "TOTAL"]
vn = ["V warps: wait weights", "V warps: loads + fma", "V warps: reduce + store"]
rays_cta0 = (b * H * H + 147) // 148
for ph in (0, 1):
    base = 32 + ph * 16
    print(f"tail phase {'AB'[ph]} (CTA 0, {rays_cta0} rays):")
    for i, n in enumerate(tn):
        print(f"  {n:26s} {100.0 * s[base + i] / max(1, s[base + 5]):6.1f}%   {s[base + i] / rays_cta0:10.0f} cyc/ray")
    for i, n in enumerate(vn):
        print(f"  {n:26s} {100.0 * s[base + 8 + i] / max(1, s[base + 5]):6.1f}%   {s[base + 8 + i] / rays_cta0:10.0f} cyc/ray")
