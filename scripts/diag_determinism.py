"""Run the 256x256 / 64-sample render several times and report which outputs differ between runs."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
quiet = len(sys.argv) > 3
b, H = 1, 256
P = int(os.environ.get("CAR_DIAG_P", "64"))
inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=1), "cuda")
z = [t.cuda() for t in synthetic.make_features(b, H, seed=1)]
m = CrossAttentionRenderer(n_view=2, npoints=P, precision=prec).cuda()
m.load_state_dict(synthetic.make_state_dict(1), strict=False); m.H = m.W = H; m.pixel_val_to_cpu = False
keys = ("rgb", "at_wt", "at_wt_max", "depth_ray", "value", "zfinal")
def run():
    taps = {"_keys": {"value", "zfinal"}}
    out = m(inp, z=z, debug_taps=taps)
    out = dict(out); out.update(taps)
    torch.cuda.synchronize()
    return {k: out[k] for k in keys}
bad = {k: 0 for k in keys}
with torch.no_grad():
    ref = run()
    for i in range(1, n + 1):
        cur = run()
        msg = []
        for k in keys:
            d = (cur[k] != ref[k])
            if d.any():
                bad[k] += 1
                idx = d.reshape(-1).nonzero()[:4, 0].tolist()
                msg.append(f"{k}: {int(d.sum())} differ (first {idx}, max abs {float((cur[k].float() - ref[k].float()).abs().max()):.3e})")
        del cur
        if not quiet:
            print(f"run {i}: " + ("identical" if not msg else "; ".join(msg)))
print("runs differing from run 0, per output:", bad, "of", n)
