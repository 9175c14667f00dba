"""Where do the milliseconds of one bench step go outside the library's kernels?  (CUDA events + host clocks)"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import synthetic, _lib, packing
from cross_attention_renderer_b200.models import CrossAttentionRenderer
b, H, P = 12, 256, 64
dev = "cuda"
inp = synthetic.to_device(synthetic.make_inputs(b, H, H, seed=100), dev)
z = [t.to(dev) for t in synthetic.make_features(b, H, seed=100)]
m = CrossAttentionRenderer(n_view=2, npoints=P, precision=sys.argv[1] if len(sys.argv) > 1 else "fp32").to(dev)
m.load_state_dict(synthetic.make_state_dict(0), strict=False); m.H = m.W = H; m.pixel_val_to_cpu = False
ev = lambda: torch.cuda.Event(enable_timing=True)
orig_launch, orig_pack = m._launch, m._packed_features
marks = {}
def timed(name, fn):
    def w(*a, **k):
        e0, e1 = ev(), ev(); t0 = time.perf_counter(); e0.record()
        r = fn(*a, **k)
        e1.record(); marks.setdefault(name, []).append((e0, e1, time.perf_counter() - t0))
        return r
    return w
m._packed_features = timed("pack_features", orig_pack)
lib = _lib.load()
orig_fwd = lib.car_render_forward
class L:  # proxy: time only the C call
    def __getattr__(self, k): return getattr(lib, k)
def step():
    m._fcache = None
    with torch.no_grad():
        return m(inp, z=z)
for _ in range(3): step()
torch.cuda.synchronize(); marks.clear()
n = 4
E0, E1 = ev(), ev(); T0 = time.perf_counter(); E0.record()
host = []
for _ in range(n):
    t0 = time.perf_counter(); step(); host.append(time.perf_counter() - t0)
E1.record(); torch.cuda.synchronize()
print(f"ms/step (events) {E0.elapsed_time(E1) / n:.2f}   wall {1e3 * (time.perf_counter() - T0) / n:.2f}")
print("host time inside forward() per step (ms):", [round(1e3 * h, 2) for h in host])
for k, v in marks.items():
    print(k, "GPU ms", [round(a.elapsed_time(b_), 2) for a, b_, _ in v], "host ms", [round(1e3 * h, 2) for _, _, h in v])
