"""GPU: the tcgen05 building blocks in isolation (operand packing, smem descriptors, UMMA issue, TMEM load):
D = A B^T against an fp64 matmul of the same operands.  bf16 mode is compared against bf16-rounded operands
(exact products, fp32 accumulation), split mode against the fp32 operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (128, 128), (32, 16), (64, 48), (128, 64), (192, 64), (256, 64)])
@pytest.mark.parametrize("split", [False, True])
def test_umma_gemm(dcf, N, K, split):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = dcf.ops.debug_umma_gemm(A, B, split=split)
    torch.cuda.synchronize()
    if split:
        ref = A.double() @ B.double().T
        tol = 2e-4      # ~K * 2^-16 relative per product, random signs
    else:
        ref = A.bfloat16().double() @ B.bfloat16().double().T
        tol = 2e-5      # exact bf16 products, fp32 accumulation order only
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, f"N={N} K={K} split={split}: rel err {err:.3e}"


def test_umma_identity_layout(dcf):
    """A = I (128x128, top-left of B^T): catches any row/column permutation in the operand layout."""
    B = torch.arange(128 * 128, dtype=torch.float32).reshape(128, 128).remainder(251.0).cuda()  # exact in bf16
    A = torch.eye(128).cuda()
    D = dcf.ops.debug_umma_gemm(A, B, split=False)
    assert torch.equal(D, B.T.contiguous())
