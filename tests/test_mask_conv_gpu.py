"""GPU parity of the union-box mask branch on this library's kernels (csrc/maskconv.cu + the bf16x3 GEMM,
lib/mask_conv.py) against fp64 torch restatements of the same ops (get_union_boxes.py:28-37 of the
reference: conv7x7/s2 + ReLU + BN + maxpool3/2/1 + conv3x3 + ReLU + BN). Layout kernels are exact
(the (hi, lo) pair reconstructs fp32 to 2^-16 relative); arithmetic kernels are checked to 3e-5 of the
output magnitude, whole-net gradients to 1e-4 relative L2 against an fp64 run on the same linear piece."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def l2err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def recon(sm):
    return sm.hi.float() + sm.lo.float()


def test_im2col7s2_matches_unfold(cuda):
    from lib import mask_conv
    torch.manual_seed(0)
    R, S = 5, 27
    m = torch.randn(R, 2, S, S, device=cuda)
    cols = F.unfold(m, 7, padding=3, stride=2)                 # [R, 2*49, 196], row = c*49 + tap
    Ho = 14
    ref = cols.view(R, 2, 49, Ho * Ho).permute(0, 3, 2, 1).reshape(R * Ho * Ho, 98)      # k = tap*2 + c
    a = mask_conv._im2col7s2(m, False)
    got = recon(a)
    assert got.shape == (R * Ho * Ho, 128)
    assert relerr(got[:, :98], ref) < 2e-5 and float(got[:, 98:].abs().max()) == 0.0
    at = mask_conv._im2col7s2(m, True)
    gt = recon(at)
    P = R * Ho * Ho
    assert gt.shape[0] == 128 and gt.shape[1] % 64 == 0
    assert relerr(gt[:98, :P], ref.t()) < 2e-5
    assert float(gt[98:].abs().max()) == 0.0 and float(gt[:, P:].abs().max()) == 0.0


@pytest.mark.parametrize("R,H,C", [(3, 7, 64), (5, 7, 256)])
def test_im2col3_nhwc_and_col2im(cuda, R, H, C):
    from lib import mask_conv
    import motifs_cabi as c
    torch.manual_seed(1)
    x = torch.randn(R, H, H, C, device=cuda)
    cols = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1)       # [R, C*9, H*H], row = c*9 + tap
    ref = cols.view(R, C, 9, H * H).permute(0, 3, 2, 1).reshape(R * H * H, 9 * C)        # k = tap*C + c
    got = recon(mask_conv._im2col3(x, False))
    assert relerr(got, ref) < 2e-5
    P = R * H * H
    gt = recon(mask_conv._im2col3(x, True))
    assert relerr(gt[:, :P], ref.t()) < 2e-5
    assert gt.shape[1] == (P + 63) // 64 * 64 and float(gt[:, P:].abs().max()) == 0.0
    # col2im is the adjoint: <im2col(x), d> == <x, col2im(d)>, and equals F.fold
    d = torch.randn(P, 9 * C, device=cuda)
    dx = torch.empty(R, H, H, C, device=cuda)
    c.check(c.load().mb200_col2im3_nhwc(c.ptr(d), R, H, H, C, c.ptr(dx), c.cur_stream()), "col2im")
    dfold = F.fold(d.view(R, H * H, 9, C).permute(0, 3, 2, 1).reshape(R, C * 9, H * H).double(), (H, H), 3, padding=1)
    assert relerr(dx, dfold.permute(0, 2, 3, 1)) < 1e-6


@pytest.mark.parametrize("P,C", [(1000, 64), (37 * 196, 256), (75264, 512)])
def test_bn_stats(cuda, P, C):
    from lib import mask_conv
    torch.manual_seed(2)
    x = (torch.randn(P, C, device=cuda) * torch.rand(C, device=cuda) * 3 + torch.randn(C, device=cuda) * 5).clamp_min(0)
    x[:, 0] = 0.0                                                # a dead channel
    x[:, 1] = 3.25                                               # a constant channel
    rm = torch.randn(C, device=cuda); rv = torch.rand(C, device=cuda) + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    mean, invstd = mask_conv._bn_stats(x, 1e-5, 0.01, rm, rv)
    xd = x.double()
    m_ref = xd.mean(0); v_ref = xd.var(0, unbiased=False)
    assert float((mean.double() - m_ref).abs().max()) < 1e-6 * float(m_ref.abs().max())
    assert relerr(invstd, 1.0 / torch.sqrt(v_ref + 1e-5)) < 1e-5
    assert relerr(rm, 0.99 * rm0.double() + 0.01 * m_ref) < 1e-6
    assert relerr(rv, 0.99 * rv0.double() + 0.01 * xd.var(0, unbiased=True)) < 1e-6


def test_bn_pool_unpool(cuda):
    import motifs_cabi as c
    torch.manual_seed(3)
    R, H, C = 7, 14, 256
    x = torch.randn(R, H, H, C, device=cuda).clamp_min(0)       # ReLU output: many exact ties at BN(0)
    mean = torch.randn(C, device=cuda) * 0.1; invstd = torch.rand(C, device=cuda) + 0.5
    gamma = torch.randn(C, device=cuda); beta = torch.randn(C, device=cuda)
    Ho = 7
    y = torch.empty(R, Ho, Ho, C, device=cuda); arg = torch.empty(R, Ho, Ho, C, dtype=torch.uint8, device=cuda)
    c.check(c.load().mb200_bn_pool3s2_nhwc(c.ptr(x), c.ptr(mean), c.ptr(invstd), c.ptr(gamma), c.ptr(beta), R, H, H, C,
                                           c.ptr(y), c.ptr(arg), c.cur_stream()), "bn_pool")
    bn = ((x - mean) * invstd * gamma + beta).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.max_pool2d(bn, 3, 2, 1)
    assert relerr(y.permute(0, 3, 1, 2), ref) < 1e-6
    gy = torch.randn(R, Ho, Ho, C, device=cuda)
    gx = torch.empty(R, H, H, C, device=cuda)
    c.check(c.load().mb200_unpool3s2_nhwc(c.ptr(gy), c.ptr(arg), R, H, H, C, c.ptr(gx), c.cur_stream()), "unpool")
    ref.backward(gy.permute(0, 3, 1, 2))
    # the same total gradient reaches each window; where values tie, ATen and this kernel both take the first maximum
    assert relerr(gx.permute(0, 3, 1, 2), bn.grad) < 1e-6


def test_layout_kernels(cuda):
    import motifs_cabi as c
    torch.manual_seed(4)
    R, HW, C = 9, 49, 512
    x = torch.randn(R, HW, C, device=cuda)
    mean = torch.randn(C, device=cuda); invstd = torch.rand(C, device=cuda) + 0.5
    gamma = torch.randn(C, device=cuda); beta = torch.randn(C, device=cuda)
    add = torch.randn(R, C, HW, device=cuda)
    out = torch.empty(R, C, HW, device=cuda)
    c.check(c.load().mb200_bn_nhwc_to_nchw(c.ptr(x), c.ptr(mean), c.ptr(invstd), c.ptr(gamma), c.ptr(beta), c.ptr(add),
                                           R, HW, C, c.ptr(out), c.cur_stream()), "bn_nhwc_to_nchw")
    ref = ((x.double() - mean.double()) * invstd.double() * gamma.double() + beta.double()).permute(0, 2, 1) + add.double()
    assert relerr(out, ref) < 1e-6
    back = torch.empty(R, HW, C, device=cuda)
    c.check(c.load().mb200_nchw_to_nhwc(c.ptr(add), R, C, HW, c.ptr(back), c.cur_stream()), "nchw_to_nhwc")
    assert torch.equal(back, add.permute(0, 2, 1).contiguous())


def test_bn_relu_backward(cuda):
    from lib import mask_conv
    torch.manual_seed(5)
    P, C = 37 * 49, 512
    z = torch.randn(P, C, device=cuda, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(C, device=cuda, dtype=torch.float64, requires_grad=True)
    beta = torch.zeros(C, device=cuda, dtype=torch.float64, requires_grad=True)
    x = z.clamp_min(0)
    y = F.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5)
    g = torch.randn(P, C, device=cuda, dtype=torch.float64)
    y.backward(g)
    xf = x.detach().float()
    mean, invstd = mask_conv._bn_stats(xf, 1e-5, 0.0, None, None)
    dz, dgamma, dbeta, dbias = mask_conv._bn_relu_backward(g.float(), xf, mean, invstd, gamma.detach().float())
    assert l2err(dz, z.grad) < 1e-5
    assert relerr(dgamma, gamma.grad) < 1e-5 and relerr(dbeta, beta.grad) < 1e-5
    assert relerr(dbias, z.grad.sum(0)) < 1e-4


@pytest.mark.parametrize("unpool", [False, True])
def test_bn_relu_backward_split_matches_unfused(cuda, unpool):
    """The fused backward (bf16-pair outputs, optional on-the-fly un-pool) against the separate kernels."""
    from lib import mask_conv
    import motifs_cabi as c
    torch.manual_seed(6)
    R, H, C = 5, 14, 256
    x = torch.randn(R, H, H, C, device=cuda).clamp_min(0)
    x2d = x.view(-1, C)
    P = x2d.size(0)
    gamma = torch.randn(C, device=cuda)
    mean, invstd = mask_conv._bn_stats(x2d, 1e-5, 0.0, None, None)
    if unpool:
        Ho = 7
        gy = torch.randn(R, Ho, Ho, C, device=cuda)
        arg = torch.randint(0, 9, (R, Ho, Ho, C), device=cuda, dtype=torch.uint8)
        arg[:, 0] = arg[:, 0].clamp_min(3); arg[:, :, 0] = (arg[:, :, 0] // 3) * 3 + (arg[:, :, 0] % 3).clamp_min(1)  # stay inside the map
        g_full = torch.empty(R, H, H, C, device=cuda)
        c.check(c.load().mb200_unpool3s2_nhwc(c.ptr(gy), c.ptr(arg), R, H, H, C, c.ptr(g_full), c.cur_stream()), "unpool")
        g_in, a_in, g2d = gy, arg, g_full.view(-1, C)
    else:
        g2d = torch.randn(P, C, device=cuda)
        g_in, a_in = g2d, None
    dz, dgamma, dbeta, dbias = mask_conv._bn_relu_backward(g2d, x2d, mean, invstd, gamma)
    t, pl, dgamma2, dbeta2, dbias2 = mask_conv._bn_relu_backward_split(g_in, a_in, x2d, mean, invstd, gamma, H, H, True)
    assert relerr(recon(pl), dz) < 2e-5
    rt = recon(t)
    assert rt.shape == (C, (P + 63) // 64 * 64)
    assert relerr(rt[:, :P], dz.t()) < 2e-5 and float(rt[:, P:].abs().max()) == 0.0
    assert relerr(dgamma2, dgamma) < 1e-6 and relerr(dbeta2, dbeta) < 1e-6 and relerr(dbias2, dbias) < 1e-5


def _make_net(dev, dim=512):
    from torch import nn
    torch.manual_seed(11)
    net = nn.Sequential(
        nn.Conv2d(2, dim // 2, kernel_size=7, stride=2, padding=3, bias=True), nn.ReLU(inplace=True),
        nn.BatchNorm2d(dim // 2, momentum=0.01), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
        nn.Conv2d(dim // 2, dim, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU(inplace=True),
        nn.BatchNorm2d(dim, momentum=0.01)).to(dev)
    with torch.no_grad():
        for m in net:
            if isinstance(m, nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    return net


def _ref_with_decisions(ref_net, masks, saved, addend):
    """fp64 forward of the branch that takes the product's own discrete decisions (ReLU masks, pool arg-max)
    from its saved tensors: gradients of a piecewise-linear net are only comparable on the same piece —
    a pool window whose two best values differ by 1e-6 routes its gradient elsewhere in fp64."""
    _, y1, _, _, arg1, _, y2, _, _, _, _, _, _ = saved
    c1, _, n1, _, c2, _, n2 = list(ref_net)
    R = masks.size(0)
    dev = masks.device
    z1 = c1(masks.double())
    x1 = z1 * (y1.view(R, 14, 14, -1) > 0).permute(0, 3, 1, 2)
    b1 = F.batch_norm(x1, None, None, n1.weight, n1.bias, True, 0.0, n1.eps)
    code = arg1.long()                                                       # [R,7,7,C1], dy*3+dx
    oy = torch.arange(7, device=dev).view(1, 7, 1, 1); ox = torch.arange(7, device=dev).view(1, 1, 7, 1)
    idx = ((2 * oy - 1 + code // 3) * 14 + (2 * ox - 1 + code % 3)).permute(0, 3, 1, 2).reshape(R, -1, 49)
    pooled = b1.flatten(2).gather(2, idx).view(R, -1, 7, 7)
    z2 = c2(pooled)
    x2 = z2 * (y2.view(R, 7, 7, -1) > 0).permute(0, 3, 1, 2)
    return F.batch_norm(x2, None, None, n2.weight, n2.bias, True, 0.0, n2.eps) + addend.detach().double()


@pytest.mark.parametrize("R,fused", [(37, True), (128, True), (37, False)])
def test_mask_conv_net_forward_backward(cuda, R, fused, monkeypatch):
    import copy
    from lib import mask_conv
    monkeypatch.setattr(mask_conv, "FUSED_BWD", fused)
    net = _make_net(cuda)
    ref_net = copy.deepcopy(net).double()
    ref_net2 = copy.deepcopy(net).double()
    assert mask_conv.supported(net)
    torch.manual_seed(R)
    masks = (torch.rand(R, 2, 27, 27, device=cuda) - 0.5) * (torch.rand(R, 2, 27, 27, device=cuda) > 0.4).float()
    addend = torch.randn(R, 512, 7, 7, device=cuda, requires_grad=True)
    gout = torch.randn(R, 512, 7, 7, device=cuda)
    # training mode: batch statistics, running-stat update, all gradients
    net.train(); ref_net.train(); ref_net2.train()
    out = mask_conv.mask_conv_net(net, masks, addend=addend)
    ref = ref_net(masks.double()) + addend.detach().double()
    assert relerr(out, ref) < 1e-4, relerr(out, ref)
    ref2 = _ref_with_decisions(ref_net2, masks, out.grad_fn.saved_tensors, addend)
    assert relerr(out, ref2) < 1e-4, relerr(out, ref2)
    out.backward(gout)
    ref.backward(gout.double())
    ref2.backward(gout.double())
    assert torch.equal(addend.grad, gout)
    for (n, p), (_, q), (_, q2) in zip(net.named_parameters(), ref_net.named_parameters(), ref_net2.named_parameters()):
        assert p.grad is not None, n
        # same piece of the piecewise-linear net: tight; independent fp64 run (its own arg-max / ReLU decisions): loose
        assert l2err(p.grad, q2.grad) < 1e-4, (n, l2err(p.grad, q2.grad))
        assert l2err(p.grad, q.grad) < 3e-2, (n, l2err(p.grad, q.grad))
    for (n, b), (_, q) in zip(net.named_buffers(), ref_net.named_buffers()):
        assert relerr(b, q) < 1e-5, n
    # eval mode: running statistics, no addend
    net.eval(); ref_net.eval()
    with torch.no_grad():
        out_e = mask_conv.mask_conv_net(net, masks)
        ref_e = ref_net(masks.double())
    assert relerr(out_e, ref_e) < 1e-4, relerr(out_e, ref_e)


def test_union_boxes_module_paths_agree(cuda, monkeypatch):
    """UnionBoxesAndFeats with MOTIFS_MASKCONV=own (default) and =cudnn give the same features and gradients."""
    import copy
    from lib import get_union_boxes as gub
    torch.manual_seed(7)
    mod = gub.UnionBoxesAndFeats(pooling_size=7, stride=16, dim=512).to(cuda).train()
    mod2 = copy.deepcopy(mod)
    B, N = 2, 12
    fmap = torch.randn(B, 512, 37, 37, device=cuda)
    xy = torch.rand(N, 2, device=cuda) * 400
    wh = torch.rand(N, 2, device=cuda) * 150 + 32
    im = torch.arange(N, device=cuda).float().div(N / B).floor()
    rois = torch.cat((im[:, None], xy, (xy + wh).clamp_max(591)), 1)
    pairs = torch.tensor([(i, j) for i in range(N) for j in range(N) if i != j and im[i] == im[j]], device=cuda)
    monkeypatch.setattr(gub, "_MASKCONV_IMPL", "own")
    a = mod(fmap, rois, pairs)
    monkeypatch.setattr(gub, "_MASKCONV_IMPL", "cudnn")
    b = mod2(fmap, rois, pairs)
    assert relerr(a, b) < 1e-4
    g = torch.randn_like(a)
    a.backward(g); b.backward(g)
    for (n, p), (_, q) in zip(mod.named_parameters(), mod2.named_parameters()):
        assert l2err(p.grad, q.grad) < 3e-2, (n, l2err(p.grad, q.grad))       # TF32 cuDNN backward + its own arg-max / ReLU decisions
