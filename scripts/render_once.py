"""Render one 256x256 / 64-sample scene a few times (short command for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b, H, P = 1, 256, 64
inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=1), "cuda")
z = [t.cuda() for t in synthetic.make_features(b, H, seed=1)]
m = CrossAttentionRenderer(n_view=2, npoints=P, precision=prec).cuda()
m.load_state_dict(synthetic.make_state_dict(1), strict=False); m.H = m.W = H; m.pixel_val_to_cpu = False
with torch.no_grad():
    for _ in range(n):
        out = m(inp, z=z)
torch.cuda.synchronize()
print("rgb checksum", float(out["rgb"].double().sum()))
