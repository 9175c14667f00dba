"""Throughput of the stand-alone tcgen05 GEMM kernels (1-CTA and CTA-pair forms)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cross_attention_renderer_b200 import _lib
lib = _lib.load()
tl = _lib.load_test()
st = torch.cuda.current_stream().cuda_stream

def split(x):
    hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

M = 148 * 128 * 16
for (N, K, nch) in [(576, 592, 3), (416, 576, 2), (256, 512, 1), (128, 128, 1)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    ah, al = split(A); wh, wl = split(W)
    C = torch.empty(M, N, device="cuda")
    for split3 in (0, 1):
        fl = 2.0 * M * N * K * (3 if split3 else 1)
        t = timeit(lambda: lib.car_gemm_umma_test(ah.data_ptr(), al.data_ptr(), wh.data_ptr(), wl.data_ptr(), None, C.data_ptr(), M, N, K, split3, 0, st))
        print(f"cg1      N={N:4d} K={K:4d} split3={split3}: {t:8.3f} ms  {fl / t / 1e9:8.1f} TF/s(mma)")
        for bk in (64, 32):
            t = timeit(lambda: tl.car_gemm_pair_test(ah.data_ptr(), al.data_ptr(), wh.data_ptr(), wl.data_ptr(), None, C.data_ptr(), None, M, N, K, nch, split3, 0, 0, bk, st))
            print(f"cg2 bk{bk} N={N:4d} K={K:4d} split3={split3}: {t:8.3f} ms  {fl / t / 1e9:8.1f} TF/s(mma)")
