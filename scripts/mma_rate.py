import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(4, dtype=torch.int64, device="cuda")
iters = 2000
for ctas in (2, 148):
    for (cg, M, N, sw) in [(1, 128, 192, 128), (1, 128, 256, 128), (1, 64, 192, 128), (1, 128, 192, 64),
                           (2, 128, 192, 128), (2, 128, 192, 64), (2, 128, 208, 64), (2, 256, 192, 128), (2, 256, 256, 128), (2, 128, 256, 64)]:
        rc = lib.car_mma_rate_test(cg, M, N, sw, iters, 2, ctas, out.data_ptr(), st)
        assert rc == 0, lib.car_last_error()
        torch.cuda.synchronize()
        cyc = int(out[0]) / (iters * 2)
        macs = M * N * 16
        print(f"ctas={ctas:3d} cg{cg} M={M:3d} N={N:3d} sw{sw:3d}: {cyc:7.1f} cyc/MMA  -> {macs / cyc / cg:7.0f} MAC/clk/SM")
