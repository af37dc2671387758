"""GPU parity of the tcgen05 bf16x3 GEMM / conv path against fp64 references of the same op.
Tolerance: the scheme keeps ~16 mantissa bits per operand with fp32 accumulation; the asserted
bound is 3e-5 of the output's max magnitude (a single-pass bf16 GEMM sits near 4e-3, tf32 near 5e-4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(params=[(0, 0), (2, 0), (2, 2)], ids=["1cta", "pair", "pair_halo"])
def kernel_mode(request):
    """Every tcgen05 kernel variant behind mb200_gemm_bf16x3 / mb200_conv3x3_bf16x3: 1-CTA tiles, CTA pairs (cta_group::2),
    and for the convolution the shared-memory halo kernel; the default picks per shape."""
    import motifs_cabi as C
    lib = C.load()
    old = (lib.mb200_gemm_set_pair_mode(request.param[0]), lib.mb200_conv_set_halo_mode(request.param[1]))
    yield request.param
    lib.mb200_gemm_set_pair_mode(old[0]); lib.mb200_conv_set_halo_mode(old[1])


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (120, 4096, 1000), (1536, 51, 4096),
                                   (300, 151, 4424), (257, 640, 712), (2000, 3072, 512), (384, 512, 1024)])
def test_gemm_bf16x3_vs_fp64(cuda, kernel_mode, M, N, K):
    from lib import tc_ops
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=cuda)
    w = torch.randn(N, K, device=cuda) / K ** 0.5
    b = torch.randn(N, device=cuda)
    ref = x.double() @ w.double().t() + b.double()
    y = tc_ops.gemm(tc_ops.split_rows(x), tc_ops.split_rows(w), bias=b)
    assert relerr(y, ref) < 3e-5, relerr(y, ref)
    yr, ys = tc_ops.gemm(tc_ops.split_rows(x), tc_ops.split_rows(w), bias=b, relu=True, want_f32=True, want_split=True)
    assert relerr(yr, ref.clamp_min(0)) < 3e-5
    rec = ys.hi[:, :N].float() + ys.lo[:, :N].float()
    assert relerr(rec, ref.clamp_min(0)) < 3e-5


def test_gemm_split_k_path(cuda):
    """Skinny M with a long K takes the split-K route (fc6 at 120 rois: K = 25088)."""
    from lib import tc_ops
    import motifs_cabi as C
    torch.manual_seed(0)
    M, N, K = 120, 1024, 25088
    assert C.load().mb200_gemm_workspace_floats(M, N, K) > 0
    x = torch.randn(M, K, device=cuda); w = torch.randn(N, K, device=cuda) / K ** 0.5
    ref = x.double() @ w.double().t()
    assert relerr(tc_ops.gemm(tc_ops.split_rows(x), tc_ops.split_rows(w)), ref) < 3e-5


def test_linear_tc_autograd(cuda):
    from lib import tc_ops
    torch.manual_seed(1)
    x = torch.randn(200, 712, device=cuda, requires_grad=True)
    lin = torch.nn.Linear(712, 300).to(cuda)
    y = tc_ops.linear_tc(x, lin.weight, lin.bias)
    g = torch.randn_like(y)
    y.backward(g)
    xd = x.detach().double().requires_grad_(True)
    wd = lin.weight.detach().double().requires_grad_(True)
    bd = lin.bias.detach().double().requires_grad_(True)
    yd = xd @ wd.t() + bd
    yd.backward(g.double())
    assert relerr(y.detach(), yd.detach()) < 3e-5
    assert relerr(x.grad, xd.grad) < 3e-5
    assert relerr(lin.weight.grad, wd.grad) < 3e-5
    assert relerr(lin.bias.grad, bd.grad) < 1e-5
    # the weight-split cache must notice an in-place update
    with torch.no_grad():
        lin.weight.add_(1.0)
    y2 = tc_ops.linear_tc(x.detach(), lin.weight, lin.bias)
    assert relerr(y2, x.detach().double() @ lin.weight.detach().double().t() + lin.bias.detach().double()) < 3e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 128), (2, 37, 37, 512, 512), (1, 74, 74, 256, 512), (2, 20, 50, 64, 64),
                                            (1, 24, 16, 64, 64), (3, 9, 70, 128, 256), (1, 5, 3, 128, 192)])
def test_conv3x3_vs_fp64(cuda, kernel_mode, B, H, W, Cin, Cout):
    from lib import tc_ops
    torch.manual_seed(B * H + W)
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).to(cuda)
    x = torch.randn(B, Cin, H, W, device=cuda)
    ref = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double(), padding=1).clamp_min(0)
    xn = x.permute(0, 2, 3, 1).contiguous().view(-1, Cin)
    xs = tc_ops.split_rows(xn)
    y, ysplit = tc_ops.conv3x3_relu((xs.hi.view(B, H, W, Cin), xs.lo.view(B, H, W, Cin)), B, H, W, Cin, conv,
                                    want_f32=True, want_split=True)
    refn = ref.permute(0, 2, 3, 1)
    assert relerr(y, refn) < 3e-5, relerr(y, refn)
    assert relerr(ysplit[0].float() + ysplit[1].float(), refn) < 3e-5


def test_maxpool_and_vgg_features_vs_torch(cuda):
    """Whole frozen VGG16 feature extractor (13 convs, 4 pools) vs torch fp64 on a small image."""
    from lib import tc_ops
    from torchvision.models.vgg import vgg16
    torch.manual_seed(3)
    feats = vgg16(weights=None).features
    del feats._modules['30']
    feats = feats.to(cuda).eval()
    x = torch.randn(2, 3, 96, 160, device=cuda)
    convs = [m for m in feats if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        ref = feats.double()(x.double()).float()
        feats.float()
        out, _ = tc_ops.vgg_features_forward(x, convs)
    assert tuple(out.shape) == (2, 6, 10, 512)
    # 13 stacked layers: north-star bar is 1e-3; the bf16x3 path should sit near 1e-4
    assert relerr(out.permute(0, 3, 1, 2), ref) < 3e-4, relerr(out.permute(0, 3, 1, 2), ref)


def test_flat_sgd_matches_torch_sgd_with_clip(cuda):
    """lib/fused_optim.FlatSGD == clip_grad_norm_ + torch.optim.SGD(momentum, weight_decay), 3 steps."""
    from lib.fused_optim import FlatSGD
    torch.manual_seed(0)
    shapes = [(33, 7), (5,), (1000, 129), (3, 3, 3)]
    a = [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    ref = torch.optim.SGD([{'params': b[:2], 'lr': 0.01}, {'params': b[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-4)
    opt = FlatSGD([(a[:2], 0.01), (a[2:], 0.1)], momentum=0.9, weight_decay=1e-4, max_norm=5.0)
    opt.zero_grad()
    for step in range(3):
        grads = [torch.randn(s, device=cuda) * (3.0 if step == 1 else 0.01) for s in shapes]   # step 1 clips
        for p, q, g in zip(a, b, grads):
            (p * g).sum().backward()          # autograd accumulates g into the flat views (and marks the parameter as used)
            q.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(b, 5.0)
        ref.step()
        opt.step()
        for p, q in zip(a, b):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), float((p - q).abs().max())
            assert float(p.grad.abs().max()) == 0.0


def test_flat_sgd_is_a_torch_optimizer(cuda):
    """The caller's recipe around the optimizer (models/train_rels.py:70, 204-205): ReduceLROnPlateau drives
    `param_groups[i]['lr']`, which the fused step reads; parameters that never received a gradient are skipped like
    torch's `grad is None` (no weight decay on them); state_dict round-trips the momentum; `defer_step` (update on a side
    stream, joined by wait_pending_updates) gives bit-identical parameters."""
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    from lib import fused_optim
    from lib.fused_optim import FlatSGD

    def build():
        torch.manual_seed(0)
        return [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in [(64, 33), (7,), (129, 5), (11,)]]

    def run(defer, steps=4, reload_at=None):
        ps = build()
        opt = FlatSGD([{'params': ps[:2], 'lr': 0.01}, {'params': ps[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-2,
                      max_norm=5.0, defer_step=defer)
        sched = ReduceLROnPlateau(opt, 'max', patience=0, factor=0.1)
        for step in range(steps):
            torch.manual_seed(100 + step)
            fused_optim.wait_pending_updates()          # what RelModel.forward does before it reads a trainable parameter
            opt.zero_grad()
            loss = sum((p * torch.randn_like(p)).sum() for p in ps[:3])         # ps[3] is never used
            loss.backward()
            opt.step()
            if step == 1:
                sched.step(0.5); sched.step(0.4)                               # no improvement -> lr * 0.1
            if reload_at == step:
                sd = opt.state_dict()
                ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
                opt2 = FlatSGD([{'params': ps2[:2], 'lr': 1.0}, {'params': ps2[2:]}], lr=1.0, momentum=0.9, weight_decay=1e-2,
                               max_norm=5.0, defer_step=defer)
                opt2.load_state_dict(sd)
                ps, opt = ps2, opt2
        fused_optim.wait_pending_updates()
        torch.cuda.synchronize()
        return ps, opt

    ps, opt = run(False)
    assert isinstance(opt, torch.optim.Optimizer)
    assert abs(opt.param_groups[0]['lr'] - 0.001) < 1e-12 and abs(opt.param_groups[1]['lr'] - 0.01) < 1e-12
    assert torch.equal(ps[3], build()[3])                                      # untouched: not even weight decay
    # torch reference with the same schedule
    qs = build()
    ref = torch.optim.SGD([{'params': qs[:2], 'lr': 0.01}, {'params': qs[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-2)
    for step in range(4):
        torch.manual_seed(100 + step)
        ref.zero_grad()
        sum((p * torch.randn_like(p)).sum() for p in qs[:3]).backward()
        torch.nn.utils.clip_grad_norm_([q for q in qs if q.grad is not None], 5.0)
        ref.step()
        if step == 1:
            for g in ref.param_groups:
                g['lr'] *= 0.1
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), float((p - q).abs().max())
    pd, _ = run(True)
    for p, q in zip(ps, pd):
        assert torch.equal(p, q)
    pr, _ = run(False, reload_at=1)
    for p, q in zip(ps, pr):
        assert torch.equal(p, q)


def test_flat_sgd_presplit_hands_over_identical_operands(cuda):
    """The fused update also writes the bf16 (hi, lo) pair of the updated parameters; for matrices with K % 64 == 0
    lib/tc_ops takes them as the GEMM operand instead of re-splitting: they must equal split_rows(param) bit for bit, stop
    being used after a foreign write + invalidate_all(), and never be offered for other shapes."""
    from lib import tc_ops
    from lib.fused_optim import FlatSGD
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(96, 128, device=cuda)); v = torch.nn.Parameter(torch.randn(40, 100, device=cuda))
    b = torch.nn.Parameter(torch.randn(128, device=cuda))
    opt = FlatSGD([([w, v, b], 0.1)], momentum=0.9, weight_decay=1e-4, max_norm=5.0)
    for step in range(2):
        opt.zero_grad()
        ((w * 2).sum() + v.pow(2).sum() + b.sum()).backward()
        opt.step()
        got = tc_ops.weight_split(w)
        assert got is w._mb200_presplit[1]                                    # the handed-over views, no kernel
        want = tc_ops.split_rows(w.detach())
        assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
        assert not hasattr(v, "_mb200_presplit") and not hasattr(b, "_mb200_presplit")      # K = 100: needs padding
        gv = tc_ops.weight_split(v); wv = tc_ops.split_rows(v.detach())
        assert torch.equal(gv.hi, wv.hi) and torch.equal(gv.lo, wv.lo)
    with torch.no_grad():
        w.data.mul_(2.0)
    tc_ops.invalidate_all()
    got = tc_ops.weight_split(w); want = tc_ops.split_rows(w.detach())
    assert got is not w._mb200_presplit[1] and torch.equal(got.hi, want.hi)


@pytest.mark.parametrize("B,H,W", [(1, 5, 7), (2, 33, 70), (1, 64, 96)])
def test_stem_conv_vs_fp64(cuda, B, H, W):
    """csrc/stem.cu: conv1_1 (3->64) + bias + ReLU, exact fp32 FMAs; output is the NHWC bf16 pair."""
    import motifs_cabi as C
    torch.manual_seed(H)
    conv = torch.nn.Conv2d(3, 64, 3, padding=1).to(cuda)
    x = torch.randn(B, 3, H, W, device=cuda)
    yh = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=cuda); yl = torch.empty_like(yh)
    C.check(C.load().mb200_conv3x3_stem_split(C.ptr(x), C.ptr(conv.weight.detach().contiguous()), C.ptr(conv.bias.detach()),
                                              B, H, W, 64, 1, C.ptr(yh), C.ptr(yl), C.cur_stream()), "stem")
    ref = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double(), padding=1).clamp_min(0)
    got = (yh.float() + yl.float()).permute(0, 3, 1, 2)
    assert relerr(got, ref) < 2e-5, relerr(got, ref)      # bounded by the bf16-pair output format (2^-17)


def test_direct_gradient_writes_match_autograd_accumulation(cuda):
    """FlatSGD parameters: weight-gradient GEMMs of linear_tc / the highway LSTM write straight into the flat
    gradient buffer (tc_ops.direct_grad_target); the result must equal autograd's accumulate path bit for bit
    (same GEMM, same inputs), over two optimizer steps."""
    from torch.nn.utils.rnn import pack_padded_sequence
    from lib import tc_ops
    from lib.fused_optim import FlatSGD
    from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import AlternatingHighwayLSTM

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = torch.nn.Linear(96, 128)
            self.rnn = AlternatingHighwayLSTM(128, 64, num_layers=2, recurrent_dropout_probability=0.0)
            self.fc2 = torch.nn.Linear(64, 10)

        def forward(self, x, lengths):
            T, B, _ = x.shape
            h = tc_ops.linear_tc(x, self.fc1.weight, self.fc1.bias).relu()
            out, _ = self.rnn(pack_padded_sequence(h, lengths))
            y = tc_ops.linear_tc(out.data, self.fc2.weight, self.fc2.bias)
            return y.pow(2).mean() + tc_ops.linear_tc(x.reshape(T * B, -1), self.fc1.weight, None).mean()   # fc1 used twice

    def run(direct):
        torch.manual_seed(3)
        net = Net().to(cuda)
        opt = FlatSGD([(list(net.parameters()), 0.05)], momentum=0.9, weight_decay=1e-4, max_norm=5.0)
        tc_ops.DIRECT_GRADS = direct
        grads = []
        try:
            for step in range(2):
                torch.manual_seed(10 + step)
                x = torch.randn(7, 4, 96, device=cuda)
                opt.zero_grad()
                net(x, torch.tensor([7, 6, 4, 2])).backward()
                grads.append([p.grad.clone() for p in net.parameters()])
                opt.step()
        finally:
            tc_ops.DIRECT_GRADS = True
        return grads, [p.detach().clone() for p in net.parameters()], net

    g_d, p_d, net = run(True)
    g_a, p_a, _ = run(False)
    assert any(len(p._mb200_direct.written) == 0 for p in net.parameters())      # states were reset by step()
    for step in range(2):
        for a, b, (n, _) in zip(g_d[step], g_a[step], net.named_parameters()):
            assert torch.equal(a, b), (step, n, float((a - b).abs().max()))
    for a, b in zip(p_d, p_a):
        assert torch.equal(a, b)
