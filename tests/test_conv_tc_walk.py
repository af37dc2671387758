"""CPU pin of the gradient formulas of lib/conv_tc.py: data gradient = the same 3x3
convolution with flipped / swapped weights, weight gradient = per-image (g^T) x (transposed im2col), through a
torch backend that implements the three primitives with the contracts of the kernel backend; checked against
torch autograd of F.conv2d in fp64."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))


@pytest.mark.parametrize("relu,bias", [(True, True), (False, True), (True, False)])
def test_conv3x3_gradients_match_autograd(relu, bias):
    from lib import conv_tc
    torch.manual_seed(0)
    B, H, W, Ci, Co = 3, 6, 9, 8, 5
    x = torch.randn(B, H, W, Ci, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.randn(Co, dtype=torch.float64, requires_grad=True) if bias else None
    y = conv_tc.conv3x3(x, w, b, relu, backend=conv_tc.TorchBackend())
    x2, w2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True) if bias else None
    ref = F.conv2d(x2.permute(0, 3, 1, 2), w2, b2, 1, 1).permute(0, 2, 3, 1)
    ref = torch.relu(ref) if relu else ref
    assert torch.allclose(y, ref, rtol=1e-12, atol=1e-12)
    g = torch.randn_like(ref)
    y.backward(g); ref.backward(g)
    assert torch.allclose(x.grad, x2.grad, rtol=1e-10, atol=1e-10)
    assert torch.allclose(w.grad, w2.grad, rtol=1e-10, atol=1e-10)
    if bias:
        assert torch.allclose(b.grad, b2.grad, rtol=1e-10, atol=1e-10)


def test_weight_matrices_layout():
    from lib import conv_tc
    w = torch.arange(2 * 3 * 9, dtype=torch.float64).reshape(2, 3, 3, 3)
    m = conv_tc.weight_matrix(w)
    assert m.shape == (2, 27) and float(m[1, (1 * 3 + 2) * 3 + 0]) == float(w[1, 0, 1, 2])
    d = conv_tc.weight_matrix_dx(w)
    assert d.shape == (3, 18) and float(d[2, (0 * 3 + 1) * 2 + 1]) == float(w[1, 2, 2, 1])


def test_vgg_features_train_walk_matches_sequential():
    from lib import conv_tc
    torch.manual_seed(1)
    cfg = [8, 8, 'M', 16, 'M', 16, 16]
    layers, convs, cin = [], [], 3
    for v in cfg:
        if v == 'M':
            layers.append(torch.nn.MaxPool2d(2, 2))
        else:
            c = torch.nn.Conv2d(cin, v, 3, padding=1).double()
            layers += [c, torch.nn.ReLU()]; convs.append(c); cin = v
    seq = torch.nn.Sequential(*layers)
    x = torch.randn(2, 3, 20, 28, dtype=torch.float64)
    ref = seq(x)
    g = torch.randn_like(ref)
    ref.backward(g)
    want = [p.grad.clone() for p in seq.parameters()]
    for p in seq.parameters():
        p.grad = None
    got = conv_tc.vgg_features_train(x, convs, cfg, backend=conv_tc.TorchBackend())
    assert torch.allclose(got.permute(0, 3, 1, 2), ref, rtol=1e-12, atol=1e-12)
    got.backward(g.permute(0, 2, 3, 1))
    for p, wgrad in zip(seq.parameters(), want):
        assert torch.allclose(p.grad, wgrad, rtol=1e-9, atol=1e-10)
