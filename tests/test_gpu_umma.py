"""GPU: the tcgen05 building blocks (descriptors, swizzle, TMEM load) against a plain fp32 matmul."""
import pytest
import torch

from mvsnerf_b200 import lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,bias", [(128, 64, False), (128, 128, False), (128, 192, True), (144, 128, True),
                                      (64, 128, True), (16, 64, False), (256, 256, False)])
def test_umma_gemm(N, K, bias):
    L = lib.load()
    g = torch.Generator().manual_seed(N * 7 + K)
    A = (torch.randn(128, K, generator=g) * 0.5).half().cuda()
    B = (torch.randn(N, K, generator=g) * 0.5).half().cuda()
    Bc = (torch.randn(N, 16, generator=g) * 0.5).half().cuda() if bias else None
    D = torch.full((128, N), float("nan"), device="cuda")
    lib.check(L.mvsn_selftest_umma(lib.ptr(A), lib.ptr(B), lib.ptr(Bc), N, K, lib.ptr(D), lib.stream_ptr()),
              "mvsn_selftest_umma")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + A[:, 16:32].float() @ Bc.float().t()
    err = (D - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
