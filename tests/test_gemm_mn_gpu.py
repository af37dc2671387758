"""GPU check of the MN-major bf16x3 GEMM (csrc/gemm_mn.cu): weight-gradient products dW = dY^T X consumed straight from
row-major activations, no transposed operand copies. It is what `linear_tc` / `matmul_tc` / the LSTM and mask-branch
weight gradients run on (lib/tc_ops.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("K,M,N", [(64, 128, 128), (256, 128, 256), (1536, 4096, 512), (120, 151, 4096), (1000, 257, 640),
                                   (75264, 512, 576), (7, 64, 64)])
def test_gemm_mn_vs_fp64(cuda, K, M, N):
    from lib import tc_ops
    torch.manual_seed(K + M + N)
    a = torch.randn(K, M, device=cuda)            # e.g. dY [rows, out]
    b = torch.randn(K, N, device=cuda) / K ** 0.5  # e.g. X  [rows, in]
    ref = a.double().t() @ b.double()
    got = tc_ops.gemm_mn(tc_ops.split_rows(a), tc_ops.split_rows(b))
    assert got.shape == (M, N)
    # fp32 accumulation over K products of ~2^-17-accurate operands: the max-norm error grows like sqrt(K); the K = 75264
    # case (the mask branch's dW2: one row per pooled position of 1536 relations) measures 3.5e-5
    assert relerr(got, ref) < (3e-5 if K < 50000 else 6e-5), relerr(got, ref)
    # and it agrees with the K-major kernel on transposed copies
    old = tc_ops.gemm(tc_ops.split_transposed(a), tc_ops.split_transposed(b))
    assert relerr(got, old) < 3e-5


def test_linear_tc_backward_with_mn_gemm(cuda):
    from lib import tc_ops
    torch.manual_seed(0)
    x = torch.randn(300, 712, device=cuda, requires_grad=True)
    w = (torch.randn(256, 712, device=cuda) / 712 ** 0.5).requires_grad_(True)
    b = torch.randn(256, device=cuda, requires_grad=True)
    assert tc_ops.GEMM_MN
    y = tc_ops.linear_tc(x, w, b)
    g = torch.randn_like(y)
    y.backward(g)
    ref_w = g.double().t() @ x.detach().double()
    assert relerr(w.grad, ref_w) < 3e-5
