import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import _lib
lib = _lib.load_test()
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(4, dtype=torch.int64, device="cuda")
iters = 2000
ctas = 2
print("single issuing thread, 2 alternating accumulators")
for (cg, M, N, sw) in [(1, 128, 32, 128), (1, 128, 64, 128), (1, 128, 128, 128), (1, 128, 256, 128), (2, 256, 256, 128), (2, 256, 64, 128)]:
    rc = lib.car_mma_rate_test(cg, M, N, sw, iters, 2, ctas, out.data_ptr(), st); assert rc == 0, lib.car_last_error()
    torch.cuda.synchronize()
    print(f"  cg{cg} M={M:3d} N={N:3d} sw{sw}: {int(out[0]) / (iters * 2):7.1f} cyc/MMA")
print("single thread, ONE accumulator (nops=1)")
for (cg, M, N, sw) in [(1, 128, 64, 128), (1, 128, 256, 128)]:
    rc = lib.car_mma_rate_test(cg, M, N, sw, iters, 1, ctas, out.data_ptr(), st); assert rc == 0
    torch.cuda.synchronize()
    print(f"  cg{cg} M={M:3d} N={N:3d} sw{sw}: {int(out[0]) / iters:7.1f} cyc/MMA")
print("k issuing warps, each its own accumulator (N=128)")
for nw in (1, 2, 4):
    for (cg, M) in [(1, 128), (2, 128), (2, 256)]:
        out.zero_()
        rc = lib.car_mma_rate_test(cg, M, 128, 128, iters, -nw, ctas, out.data_ptr(), st); assert rc == 0, lib.car_last_error()
        torch.cuda.synchronize()
        t = max(int(x) for x in out[:nw])
        print(f"  warps={nw} cg{cg} M={M}: {t / iters:7.1f} cyc per MMA-per-warp -> aggregate {t / (iters * nw):6.1f} cyc/MMA")
