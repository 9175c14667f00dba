"""Per-step device time of 20 consecutive bench steps in a fresh process (does the step time settle?)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic
from cross_attention_renderer_b200.models import CrossAttentionRenderer
b, H, P = 12, 256, 64
dev = "cuda"
inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=100), dev)
z = [t.to(dev) for t in synthetic.make_features(b, H, seed=100)]
m = CrossAttentionRenderer(n_view=2, npoints=P, precision="fp32").to(dev)
m.load_state_dict(synthetic.make_state_dict(0), strict=False); m.H = m.W = H; m.pixel_val_to_cpu = False
evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
evs[0].record()
for i in range(20):
    m._fcache = None
    with torch.no_grad():
        m(inp, z=z)
    evs[i + 1].record()
torch.cuda.synchronize()
print("ms per step:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(20)])
if len(sys.argv) > 1:
    import ctypes as C
    from cross_attention_renderer_b200 import _lib
    lib = _lib.load()
    nst = len(_lib.STAGES)
    ms_arr, ln_arr = (C.c_float * nst)(), (C.c_int * nst)()
    lib.car_profile_begin()
    evs[0].record()
    for i in range(10):
        m._fcache = None
        with torch.no_grad():
            m(inp, z=z)
        evs[i + 1].record()
    torch.cuda.synchronize()
    lib.car_profile_end(ms_arr, ln_arr, nst)
    print("with stage events, ms per step:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(10)], "sum of stages/step", round(sum(ms_arr) / 10, 1))
