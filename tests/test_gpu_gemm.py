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


def _pair_call(lib, A, W, bias, nch, split3, relu=1, max_pairs=0, dump=False, bk=64):
    M, K = A.shape
    N = W.shape[0]
    ah, al = split(A)
    wh, wl = split(W)
    C = torch.full((M, N), float("nan"), device="cuda")
    npairs = min(74, (M + 127) // 128) if not max_pairs else max_pairs
    D = torch.full((npairs * 2, 128, N // 2), float("nan"), device="cuda") if dump else None
    rc = lib.car_gemm_pair_test(ah.data_ptr(), al.data_ptr() if split3 else None, wh.data_ptr(),
                                wl.data_ptr() if split3 else None, bias.data_ptr() if bias is not None else None,
                                C.data_ptr(), D.data_ptr() if dump else None, M, N, K, nch, split3, relu,
                                max_pairs, bk, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.car_last_error()
    torch.cuda.synchronize()
    return C, D, (ah, wh)


def test_pair_gemm_layout_probe():
    """D[m][n] = 256*m + n makes the TMEM image self-describing: print where rows/columns of the
    cta_group::2 accumulator live (diagnostic for the 2x2 datapath layout assumption)."""
    lib = _lib.load_test()
    M, N, K = 128, 64, 16
    A = torch.zeros(M, K); W = torch.zeros(N, K)
    A[:, 0] = torch.arange(M).float(); A[:, 1] = 1.0
    W[:, 0] = 256.0; W[:, 1] = torch.arange(N).float()
    C, D, _ = _pair_call(lib, A.cuda(), W.cuda(), None, 1, 0, relu=0, max_pairs=1, dump=True)
    D = D.cpu()
    m_of = torch.div(D, 256, rounding_mode="floor")
    n_of = D - 256 * m_of
    for cta in range(2):
        for ln in (0, 1, 31, 32, 63, 64, 65, 127):
            print(f"cta{cta} lane{ln:3d}: m={m_of[cta, ln, :3].tolist()} n(first3)={n_of[cta, ln, :3].tolist()} n(last)={n_of[cta, ln, -1].item()}")
    ref = (A @ W.T)
    assert torch.equal(C.cpu(), ref), "assumed 2x2 layout / operand split is wrong (see probe above)"


@pytest.mark.parametrize("M,N,K,nch", [(128, 192, 64, 1), (300, 576, 592, 3), (5000, 576, 592, 3),
                                       (1000, 416, 576, 2), (1000, 128, 128, 1), (20000, 128, 16, 1)])
@pytest.mark.parametrize("split3", [0, 1])
@pytest.mark.parametrize("bk", [64, 32])
def test_pair_gemm(M, N, K, nch, split3, bk):
    lib = _lib.load_test()
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C, _, (ah, wh) = _pair_call(lib, A, W, bias, nch, split3, bk=bk)
    if split3:
        ref = torch.relu(A.double() @ W.double().T + bias.double()); tol = 3e-5
    else:
        ref = torch.relu(ah.double() @ wh.double().T + bias.double()); tol = 2e-5
    assert torch.isfinite(C).all()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, err
