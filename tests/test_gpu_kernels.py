"""-m gpu: kernel-level checks of the C-ABI GEMMs against torch fp32 / fp64 matmul on the same inputs.

spv_gemm (fp32 SIMT): relative error <= 1e-5.  spv_tc_gemm (bf16 tcgen05): inputs are rounded to bf16 on both sides,
products accumulate in fp32, so the comparison against torch on the SAME bf16-rounded inputs is <= 1e-4 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(70, 130, 45), (512, 256, 1000), (33, 7, 300)])
def test_simt_gemm(ta, tb, M, N, K):
    from spvipes_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((K, M) if ta else (M, K), generator=g, device="cuda")
    B = torch.randn((N, K) if tb else (K, N), generator=g, device="cuda")
    bias = torch.randn(N, generator=g, device="cuda")
    C = torch.empty(M, N, device="cuda")
    ws = torch.empty(4 * M * N, device="cuda")
    for splits in (1, 4):
        L.check(lib.spv_gemm(0, ta, 0, tb, A.data_ptr(), A.stride(0), None, B.data_ptr(), B.stride(0), None, C.data_ptr(), N, M, N,
                             K, 1, 0, 0, 0, bias.data_ptr(), 0, 1, 0, splits, ws.data_ptr(), _stream()), "spv_gemm")
        torch.cuda.synchronize()
        want = torch.relu((A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double())
        err = float((C.double() - want).abs().max() / want.abs().max())
        assert err < 1e-5, (splits, err)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (512, 256, 5000), (200, 291, 512), (5000, 40, 512), (96, 64, 136)])
def test_tc_gemm(a_mn, b_mn, M, N, K):
    from spvipes_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(2)
    r8 = lambda x: (x + 7) // 8 * 8
    A = torch.zeros((K, r8(M)) if a_mn else (M, r8(K)), device="cuda", dtype=torch.bfloat16)
    B = torch.zeros((K, r8(N)) if b_mn else (N, r8(K)), device="cuda", dtype=torch.bfloat16)
    if a_mn:
        A[:, :M] = torch.randn(K, M, generator=g, device="cuda").bfloat16()
    else:
        A[:, :K] = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    if b_mn:
        B[:, :N] = torch.randn(K, N, generator=g, device="cuda").bfloat16()
    else:
        B[:, :K] = torch.randn(N, K, generator=g, device="cuda").bfloat16()
    Af = (A[:, :M].t() if a_mn else A[:, :K]).double()
    Bf = (B[:, :N].t() if b_mn else B[:, :K]).double()
    bias = torch.randn(N, generator=g, device="cuda")
    want = Af @ Bf.t() + bias.double()
    ws = torch.empty(8 * M * N, device="cuda")
    for splits in (1, 3):
        C = torch.full((M, N), float("nan"), device="cuda")
        L.check(lib.spv_tc_gemm(a_mn, b_mn, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), N, M, N, K,
                                bias.data_ptr(), 0, 0, splits, ws.data_ptr(), _stream()), "spv_tc_gemm")
        torch.cuda.synchronize()
        err = float((C.double() - want).abs().max() / want.abs().max())
        assert err < 1e-4, (splits, err)
