"""The recurrent path's dense contraction (mava_gemm): fp32 SIMT kernel and bf16 tcgen05 kernel in
every operand arrangement, epilogue and accumulation mode, against torch float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float64)


CASES = [
    # M, N, K, ta, tb
    (300, 128, 70, 0, 0),     # X W (odd K, unaligned rows)
    (257, 384, 128, 0, 0),    # E Wi
    (131, 5, 128, 0, 0),      # head, tiny N
    (260, 1, 24, 0, 0),       # critic head
    (200, 128, 384, 0, 1),    # dG Wh^T
    (128, 384, 1000, 1, 0),   # Hin^T dG (weight gradient)
    (213, 128, 900, 1, 0),    # X^T dE1, odd M
    (24, 13, 515, 1, 0),      # P^T dlogits
    (150, 96, 200, 1, 1),     # both transposed
]


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("M,N,K,ta,tb", CASES)
def test_gemm_matches_torch(lib_built, use_tc, M, N, K, ta, tb):
    from mava_b200 import native

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref_mask = torch.randn(M, N, generator=g)
    Ad, Bd = A.to(dev), B.to(dev)
    A64, B64 = (A.double(), B.double()) if not use_tc else (_bf16_round(A), _bf16_round(B))
    prod = (A64.T if ta else A64) @ (B64.T if tb else B64)
    tol = dict(rtol=1e-5, atol=1e-4) if not use_tc else dict(rtol=1e-4, atol=1e-3)

    # store + bias + relu
    C = torch.full((M, N), 7.0, device=dev)
    native.gemm(use_tc, Ad, ta, Bd, tb, C, M, N, K, bias=bias.to(dev), relu=True)
    np.testing.assert_allclose(C.cpu().numpy(), torch.relu(prod + bias).numpy(), **tol)
    # accumulate + relu' mask
    C0 = torch.randn(M, N, generator=g)
    C = C0.clone().to(dev)
    native.gemm(use_tc, Ad, ta, Bd, tb, C, M, N, K, relu_ref=ref_mask.to(dev), mode=1)
    want = C0.double() + torch.where(ref_mask > 0, prod, torch.zeros_like(prod))
    np.testing.assert_allclose(C.cpu().numpy(), want.numpy(), **tol)
    # in-place mask (C aliases relu_ref), as the backward pass uses it
    C = ref_mask.clone().to(dev)
    native.gemm(use_tc, Ad, ta, Bd, tb, C, M, N, K, relu_ref=C)
    np.testing.assert_allclose(C.cpu().numpy(),
                               torch.where(ref_mask > 0, prod, torch.zeros_like(prod)).numpy(), **tol)
    # split-K with atomics
    C = torch.zeros(M, N, device=dev)
    native.gemm(use_tc, Ad, ta, Bd, tb, C, M, N, K, mode=2, k_splits=3)
    np.testing.assert_allclose(C.cpu().numpy(), prod.numpy(), rtol=tol["rtol"] * 5,
                               atol=tol["atol"] * 5)


def test_tc_gemm_strided_operands(lib_built):
    """Leading dimensions larger than the row length (the gate stash is read with ld = 4H)."""
    from mava_b200 import native

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    big = torch.randn(200, 512, generator=g).to(dev)
    W = torch.randn(128, 384, generator=g).to(dev)
    A = big[:, :384]  # 200 x 384 with ld 512
    for use_tc in (0, 1):
        C = torch.zeros(200, 128, device=dev)
        native.gemm(use_tc, A, 0, W, 1, C, 200, 128, 384)
        a64 = _bf16_round(A.cpu()) if use_tc else A.cpu().double()
        w64 = _bf16_round(W.cpu()) if use_tc else W.cpu().double()
        np.testing.assert_allclose(C.cpu().numpy(), (a64 @ w64.T).numpy(), rtol=1e-4, atol=2e-3)
