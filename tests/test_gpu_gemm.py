"""tcgen05 GEMM (car_gemm_umma.cu) against torch on the same bf16-rounded operands."""
import pytest
import torch

from cross_attention_renderer_b200 import _lib

pytestmark = pytest.mark.gpu


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 128, 16), (1000, 576, 592), (4096, 288, 576),
                                   (77, 128, 288), (40000, 576, 592), (33000, 128, 128)])
@pytest.mark.parametrize("split3", [0, 1])
def test_gemm_umma(M, N, K, split3):
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ah, al = split(A)
    wh, wl = split(W)
    C = torch.full((M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.car_gemm_umma_test(ah.data_ptr(), al.data_ptr() if split3 else None, wh.data_ptr(),
                                wl.data_ptr() if split3 else None, bias.data_ptr(), C.data_ptr(), M, N, K,
                                split3, 1, st)
    assert rc == 0, lib.car_last_error()
    torch.cuda.synchronize()
    if split3:
        ref = torch.relu(A.double() @ W.double().T + bias.double())
        tol = 3e-5
    else:
        ref = torch.relu(ah.double() @ wh.double().T + bias.double())
        tol = 2e-5          # fp32 accumulation of exact bf16 products
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert torch.isfinite(C).all()
    assert err < tol, err
