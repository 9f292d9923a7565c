"""GPU: the two tcgen05 GEMM shapes behind cf_fusion_bwd in isolation, against fp64 matmuls of the same fp32 operands.
Split-bf16 arithmetic (hi*hi + hi*lo + lo*hi, fp32 accumulate): ~2^-16 relative per product, tolerance 1e-4 of the
largest entry.  Shapes are the ones K-4b uses (C = 32...256 against C or Ci = 128), ragged row counts, row counts that
live in device memory, every epilogue."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _err(a, ref):
    return (a.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


@pytest.mark.parametrize("Kd,N", [(32, 32), (64, 64), (96, 96), (128, 128), (32, 128), (64, 128), (192, 128), (256, 128)])
@pytest.mark.parametrize("transpose", [False, True])
def test_bwd_gemm_nn_store(dcf, Kd, N, transpose):
    g = torch.Generator().manual_seed(Kd * 7 + N)
    R = 128 * 5 + 37
    X = torch.randn(R, Kd, generator=g).cuda()
    Wm = (torch.randn(Kd, N, generator=g) if transpose else torch.randn(N, Kd, generator=g)).cuda()
    out = dcf.ops.debug_bwd_gemm_nn(X, Wm, transpose=transpose)
    ref = X.double() @ (Wm.double() if transpose else Wm.double().T)
    assert _err(out, ref) < TOL


def test_bwd_gemm_nn_epilogues_and_device_row_count(dcf):
    g = torch.Generator().manual_seed(3)
    R, Kd, N, live = 3000, 64, 64, 1777
    X = torch.randn(R, Kd, generator=g).cuda()
    Wm = torch.randn(N, Kd, generator=g).cuda()
    ref = X.double() @ Wm.double().T
    cnt = torch.tensor([live], dtype=torch.int32, device="cuda")
    # accumulate, rows beyond the device-side count untouched
    base = torch.randn(R, N, generator=g).cuda()
    out = dcf.ops.debug_bwd_gemm_nn(X, Wm, epi=1, out=base.clone(), row_count=cnt)
    assert _err(out[:live], base[:live].double() + ref[:live]) < TOL
    assert torch.equal(out[live:], base[live:])
    # relu(. + bias)
    bias = torch.randn(N, generator=g).cuda()
    out = dcf.ops.debug_bwd_gemm_nn(X, Wm, epi=2, aux=bias)
    assert _err(out, torch.relu(ref + bias.double())) < TOL
    # mask read from the output buffer itself (the way dA overwrites H1)
    buf = torch.randn(R, N, generator=g).cuda()
    mask = buf > 0
    out = dcf.ops.debug_bwd_gemm_nn(X, Wm, epi=3, aux=buf, out=buf)
    assert _err(out, ref * mask) < TOL
    assert (out[~mask] == 0).all()


@pytest.mark.parametrize("M,N,n2,bias", [(32, 32, 0, True), (64, 64, 0, True), (128, 128, 0, True), (192, 192, 0, True),
                                          (256, 256, 0, False), (32, 128, 0, True), (256, 128, 3, True), (64, 0, 3, False)])
def test_bwd_gemm_tn(dcf, M, N, n2, bias):
    g = torch.Generator().manual_seed(M * 5 + N + n2)
    R = 64 * 150 + 21
    X = torch.randn(R, M, generator=g).cuda()
    Y = torch.randn(R, N, generator=g).cuda() if N else None
    Y2 = torch.randn(R, n2, generator=g).cuda() if n2 else None
    wcol = torch.rand(R, generator=g).cuda() if bias and M == 64 else None
    dW, db = dcf.ops.debug_bwd_gemm_tn(X, Y, Y2, wcol=wcol, bias=bias)
    cat = torch.cat([t.double() for t in (Y, Y2) if t is not None], 1)
    ref = X.double().T @ cat
    assert _err(dW, ref) < TOL
    if bias:
        w = wcol.double() if wcol is not None else torch.ones(R, dtype=torch.float64, device="cuda")
        assert _err(db, X.double().T @ w) < TOL


def test_bwd_gemm_tn_device_row_count(dcf):
    g = torch.Generator().manual_seed(9)
    R, live = 20000, 4321
    X = torch.randn(R, 64, generator=g).cuda()
    Y = torch.randn(R, 64, generator=g).cuda()
    cnt = torch.tensor([live], dtype=torch.int32, device="cuda")
    dW, db = dcf.ops.debug_bwd_gemm_tn(X, Y, bias=True, row_count=cnt)
    assert _err(dW, X[:live].double().T @ Y[:live].double()) < TOL
    assert _err(db, X[:live].double().sum(0)) < TOL
