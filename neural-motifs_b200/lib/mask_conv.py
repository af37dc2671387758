"""The mask branch of UnionBoxesAndFeats on this library's own kernels.

Reference: `self.conv` of lib/get_union_boxes.py:28-37 —
    Conv2d(2, dim/2, 7, stride 2, pad 3) -> ReLU -> BatchNorm2d -> MaxPool2d(3, 2, 1)
    -> Conv2d(dim/2, dim, 3, pad 1) -> ReLU -> BatchNorm2d
applied to the [R,2,27,27] masks of draw_union_boxes and added to the union RoIAlign features (:53).

Both convolutions are explicit-im2col x weight products on the tcgen05 bf16x3 GEMM (csrc/gemm_tc.cu);
BatchNorm statistics / apply, the pool, the layout changes and the whole backward are the NHWC streaming
kernels of csrc/maskconv.cu. Nothing here falls back to cuDNN: `mask_conv_net` needs the CUDA library.
"""
import os

import torch
from torch.autograd import Function

import motifs_cabi as _c
from lib import tc_ops
from lib.tc_ops import SplitMat, gemm, gemm_mn, split_rows, split_transposed, _cached, _round_up


def _lib_call(name, dev, *args):
    with torch.cuda.device(dev):
        _c.check(getattr(_c.load(), name)(*args, _c.cur_stream()), name)


def _im2col7s2(masks, transposed):
    R, _, S, _ = masks.shape
    Ho = (S - 1) // 2 + 1
    P = R * Ho * Ho
    dev = masks.device
    if transposed:
        Pp = _round_up(P, 64)
        hi = torch.empty(128, Pp, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
        _lib_call("mb200_im2col7s2_split", dev, _c.ptr(masks), R, S, 1, Pp, _c.ptr(hi), _c.ptr(lo))
        return SplitMat(hi, lo, 128, P, Pp)
    hi = torch.empty(P, 128, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
    _lib_call("mb200_im2col7s2_split", dev, _c.ptr(masks), R, S, 0, 0, _c.ptr(hi), _c.ptr(lo))
    return SplitMat(hi, lo, P, 98, 128)


def _im2col3(x_nhwc, transposed):
    R, H, W, C = x_nhwc.shape
    P = R * H * W
    dev = x_nhwc.device
    if transposed:
        Pp = _round_up(P, 64)
        hi = torch.empty(9 * C, Pp, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
        _lib_call("mb200_im2col3_nhwc_split", dev, _c.ptr(x_nhwc), R, H, W, C, 1, Pp, _c.ptr(hi), _c.ptr(lo))
        return SplitMat(hi, lo, 9 * C, P, Pp)
    hi = torch.empty(P, 9 * C, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
    _lib_call("mb200_im2col3_nhwc_split", dev, _c.ptr(x_nhwc), R, H, W, C, 0, 0, _c.ptr(hi), _c.ptr(lo))
    return SplitMat(hi, lo, P, 9 * C, 9 * C)


def _w_stem(w):       # [O,2,7,7] -> B operand [O, 128], k = (ky*7+kx)*2 + c
    return _cached(w, "mask_stem", lambda t: split_rows(t.permute(0, 2, 3, 1).reshape(t.size(0), -1).contiguous()))


def _w3_mat(t):       # [O,I,3,3] -> [O, 9I], k = (ky*3+kx)*I + i
    return t.permute(0, 2, 3, 1).reshape(t.size(0), -1).contiguous()


def _w3(w):           # B operand of the forward product
    return _cached(w, "mask_w3", lambda t: split_rows(_w3_mat(t)))


def _w3_t(w):         # [9I, O]: B operand of dcol = dz @ Wmat
    return _cached(w, "mask_w3_t", lambda t: split_transposed(_w3_mat(t)))


def _bn_stats(x2d, eps, momentum, run_mean, run_var):
    P, C = x2d.shape
    dev = x2d.device
    sums = torch.empty(4 * C, dtype=torch.float64, device=dev)
    mean = torch.empty(C, dtype=torch.float32, device=dev)
    invstd = torch.empty(C, dtype=torch.float32, device=dev)
    _lib_call("mb200_bn_stats", dev, _c.ptr(x2d), P, C, float(eps), float(momentum), _c.ptr(sums), _c.ptr(mean),
              _c.ptr(invstd), _c.ptr(run_mean), _c.ptr(run_var))
    return mean, invstd


def _bn_relu_backward(g2d, x2d, mean, invstd, gamma):
    P, C = x2d.shape
    dev = x2d.device
    sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dbias = torch.empty(C, dtype=torch.float64, device=dev)
    dz = torch.empty(P, C, dtype=torch.float32, device=dev)
    _lib_call("mb200_bn_relu_backward", dev, _c.ptr(g2d), _c.ptr(x2d), _c.ptr(mean), _c.ptr(invstd), _c.ptr(gamma), P, C,
              _c.ptr(sums), _c.ptr(dz), _c.ptr(dbias))
    return dz, sums[C:].float(), sums[:C].float(), dbias.float()      # dz, dgamma, dbeta, dconv_bias


# "1": BN/ReLU backward emits the GEMM operands (bf16 pairs, transposed / plain) directly and routes the pooled
# gradient through the max-pool on the fly; "0": separate un-pool, fp32 dz, split kernels (kept for A/B runs).
FUSED_BWD = os.environ.get("MOTIFS_MASKCONV_FUSED_BWD", "0") == "1"


def _bn_relu_backward_split(g, arg, x2d, mean, invstd, gamma, H, W, want_plain):
    """mb200_bn_relu_backward_split: returns (dz^T SplitMat [C, P], dz SplitMat [P, C] or None, dgamma, dbeta, dbias).
    arg is None: g is [P,C]; else g is the pooled gradient [R,Ho,Wo,C] and arg its arg-max codes."""
    P, C = x2d.shape
    dev = x2d.device
    Pp = _round_up(P, 64)
    sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
    dbias = torch.empty(C, dtype=torch.float64, device=dev)
    t_hi = torch.empty(C, Pp, dtype=torch.bfloat16, device=dev); t_lo = torch.empty_like(t_hi)
    p_hi = p_lo = None
    if want_plain:
        p_hi = torch.empty(P, C, dtype=torch.bfloat16, device=dev); p_lo = torch.empty_like(p_hi)
    _lib_call("mb200_bn_relu_backward_split", dev, _c.ptr(g), _c.ptr(arg), _c.ptr(x2d), _c.ptr(mean), _c.ptr(invstd),
              _c.ptr(gamma), P, Pp, C, H, W, _c.ptr(sums), _c.ptr(t_hi), _c.ptr(t_lo), _c.ptr(p_hi), _c.ptr(p_lo),
              _c.ptr(dbias))
    plain = SplitMat(p_hi, p_lo, P, C, C) if want_plain else None
    return SplitMat(t_hi, t_lo, C, P, Pp), plain, sums[C:].float(), sums[:C].float(), dbias.float()


class _MaskConvNet(Function):
    """out[R,C2,7,7] = addend + BN2(ReLU(conv3x3(pool(BN1(ReLU(conv7x7s2(masks)))))))."""

    @staticmethod
    def forward(ctx, masks, addend, w1, b1, g1, be1, w2, b2, g2, be2, bn1, bn2, training):
        _c.require_cuda(masks, w1, w2)
        masks = masks.detach().contiguous().float()
        dev = masks.device
        R, _, S, _ = masks.shape
        C1, C2 = w1.size(0), w2.size(0)
        H1 = (S - 1) // 2 + 1
        H2 = (H1 - 1) // 2 + 1
        # conv1 (+bias, ReLU) -> y1 NHWC [R*H1*H1, C1]
        col1 = _im2col7s2(masks, False)
        y1 = gemm(col1, _w_stem(w1), bias=b1.detach(), relu=True)
        if training:
            mean1, inv1 = _bn_stats(y1, bn1.eps, bn1.momentum, bn1.running_mean, bn1.running_var)
        else:
            mean1, inv1 = bn1.running_mean, torch.rsqrt(bn1.running_var + bn1.eps)
        p1 = torch.empty(R, H2, H2, C1, dtype=torch.float32, device=dev)
        arg1 = torch.empty(R, H2, H2, C1, dtype=torch.uint8, device=dev)
        _lib_call("mb200_bn_pool3s2_nhwc", dev, _c.ptr(y1), _c.ptr(mean1), _c.ptr(inv1), _c.ptr(g1.detach()),
                  _c.ptr(be1.detach()), R, H1, H1, C1, _c.ptr(p1), _c.ptr(arg1))
        # conv2 (+bias, ReLU) -> y2 NHWC [R*H2*H2, C2]
        col2 = _im2col3(p1, False)
        y2 = gemm(col2, _w3(w2), bias=b2.detach(), relu=True)
        if training:
            mean2, inv2 = _bn_stats(y2, bn2.eps, bn2.momentum, bn2.running_mean, bn2.running_var)
        else:
            mean2, inv2 = bn2.running_mean, torch.rsqrt(bn2.running_var + bn2.eps)
        out = torch.empty(R, C2, H2, H2, dtype=torch.float32, device=dev)
        add = addend.detach().contiguous() if addend is not None else None
        _lib_call("mb200_bn_nhwc_to_nchw", dev, _c.ptr(y2), _c.ptr(mean2), _c.ptr(inv2), _c.ptr(g2.detach()),
                  _c.ptr(be2.detach()), _c.ptr(add), R, H2 * H2, C2, _c.ptr(out))
        ctx.training = training
        ctx.dims = (R, S, H1, H2, C1, C2)
        ctx.has_addend = addend is not None
        ctx.save_for_backward(masks, y1, mean1, inv1, arg1, p1, y2, mean2, inv2, w1, g1, w2, g2)
        # the plain im2col pairs are the B operands of the weight-gradient GEMMs on the MN-major kernel (dW = dz^T col):
        # keeping them (0.85 GB at 1536 relations) removes the transposed im2col passes of backward
        ctx.cols = (col1, col2) if (training and tc_ops.GEMM_MN and not FUSED_BWD) else None
        return out

    @staticmethod
    def backward(ctx, g):
        if not ctx.training:
            raise NotImplementedError("mask_conv_net: backward needs training-mode BatchNorm (batch statistics)")
        masks, y1, mean1, inv1, arg1, p1, y2, mean2, inv2, w1, g1, w2, g2 = ctx.saved_tensors
        R, S, H1, H2, C1, C2 = ctx.dims
        dev = g.device
        g = g.contiguous()
        # BN2 / ReLU backward in NHWC
        g_nhwc = torch.empty(R * H2 * H2, C2, dtype=torch.float32, device=dev)
        _lib_call("mb200_nchw_to_nhwc", dev, _c.ptr(g), R, C2, H2 * H2, _c.ptr(g_nhwc))
        if FUSED_BWD:
            dz2_t, dz2_p, dg2, dbe2, db2 = _bn_relu_backward_split(g_nhwc, None, y2, mean2, inv2, g2.detach(), H2, H2, True)
        else:
            dz2, dg2, dbe2, db2 = _bn_relu_backward(g_nhwc, y2, mean2, inv2, g2.detach())
            dz2_p = split_rows(dz2)
            dz2_t = split_transposed(dz2) if ctx.cols is None else None
            del dz2
        del g_nhwc
        # conv2: dW2 = dz2^T @ im2col(p1);  dp1 = col2im(dz2 @ W2mat)
        if ctx.cols is not None:
            dw2 = gemm_mn(dz2_p, ctx.cols[1])                                         # [C2, 9*C1], no transposed operands
        else:
            dw2 = gemm(dz2_t, _im2col3(p1, True))                                     # [C2, 9*C1]
        dw2 = dw2.view(C2, 3, 3, C1).permute(0, 3, 1, 2).contiguous()
        dcol = gemm(dz2_p, _w3_t(w2))                                                 # [P2, 9*C1]
        del dz2_t, dz2_p
        dp1 = torch.empty(R, H2, H2, C1, dtype=torch.float32, device=dev)
        _lib_call("mb200_col2im3_nhwc", dev, _c.ptr(dcol), R, H2, H2, C1, _c.ptr(dp1))
        del dcol
        # pool / BN1 / ReLU backward
        if FUSED_BWD:
            dz1_t, _, dg1, dbe1, db1 = _bn_relu_backward_split(dp1, arg1, y1, mean1, inv1, g1.detach(), H1, H1, False)
        else:
            dbn1 = torch.empty(R * H1 * H1, C1, dtype=torch.float32, device=dev)
            _lib_call("mb200_unpool3s2_nhwc", dev, _c.ptr(dp1), _c.ptr(arg1), R, H1, H1, C1, _c.ptr(dbn1))
            dz1, dg1, dbe1, db1 = _bn_relu_backward(dbn1, y1, mean1, inv1, g1.detach())
            del dbn1
            dz1_t = split_transposed(dz1) if ctx.cols is None else split_rows(dz1)
            del dz1
        # conv1: dW1 = dz1^T @ im2col(masks) (the masks themselves need no gradient)
        if ctx.cols is not None:
            dw1 = gemm_mn(dz1_t, ctx.cols[0])                                         # [C1, 128]
            ctx.cols = None
        else:
            dw1 = gemm(dz1_t, _im2col7s2(masks, True))                                # [C1, 128]
        dw1 = dw1[:, :98].reshape(C1, 7, 7, 2).permute(0, 3, 1, 2).contiguous()
        return (None, g if ctx.has_addend else None, dw1, db1, dg1, dbe1, dw2, db2, dg2, dbe2, None, None, None)


def supported(conv_seq):
    """True when `conv_seq` is exactly the reference's mask branch (get_union_boxes.py:28-37)."""
    from torch import nn
    m = list(conv_seq)
    if len(m) != 7:
        return False
    c1, r1, n1, pl, c2, r2, n2 = m
    ok = isinstance(c1, nn.Conv2d) and c1.kernel_size == (7, 7) and c1.stride == (2, 2) and c1.padding == (3, 3) \
        and c1.in_channels == 2 and c1.bias is not None
    ok = ok and isinstance(r1, nn.ReLU) and isinstance(n1, nn.BatchNorm2d) and n1.affine and n1.track_running_stats
    ok = ok and isinstance(pl, nn.MaxPool2d) and pl.kernel_size == 3 and pl.stride == 2 and pl.padding == 1
    ok = ok and isinstance(c2, nn.Conv2d) and c2.kernel_size == (3, 3) and c2.stride == (1, 1) and c2.padding == (1, 1) \
        and c2.bias is not None and c2.in_channels % 32 == 0 and c2.in_channels == c1.out_channels \
        and c2.out_channels % 64 == 0
    ok = ok and isinstance(r2, nn.ReLU) and isinstance(n2, nn.BatchNorm2d) and n2.affine and n2.track_running_stats
    return bool(ok and n1.momentum is not None and n2.momentum is not None)


def mask_conv_net(conv_seq, masks, addend=None):
    """conv_seq(masks) (+ addend) with `conv_seq` the nn.Sequential of get_union_boxes.py:28-37."""
    c1, _, n1, _, c2, _, n2 = list(conv_seq)
    training = conv_seq.training
    out = _MaskConvNet.apply(masks, addend, c1.weight, c1.bias, n1.weight, n1.bias, c2.weight, c2.bias,
                             n2.weight, n2.bias, n1, n2, training)
    if training:
        with torch.no_grad():
            n1.num_batches_tracked += 1
            n2.num_batches_tracked += 1
    return out
