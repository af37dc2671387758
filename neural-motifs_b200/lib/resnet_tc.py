"""ResNet-101 conv1..layer3 (C4, stride 16, 1024 channels) of the detector's `use_resnet=True` branch —
`ObjectDetector.feature_map` of the reference, lib/object_detector.py:119-127, over torchvision's
`resnet101` modules (`load_resnet`, :615-620) — walked layer by layer on this library's kernels.

STATUS (SURVEY.md §8a row a1'): the graph walk below is pinned on the CPU against torchvision's own forward through a
test backend (tests/test_resnet_walk.py); the kernel backend is held to fp64 and to the oracle's detector on a B200 by
tests/test_resnet_gpu.py (green since round 2).

Layout: activations are NHWC fp32 [B,H,W,C] between operations. A backend supplies five operations:
    stem(x_nchw, conv)            7x7 / stride 2 / pad 3, 3 -> 64
    conv1x1(x, conv)              stride 1 or 2 (stride 2 = row/column subsampling, then a GEMM)
    conv3x3(x, conv)              pad 1, stride 1 or 2 (stride 2 = the stride-1 result subsampled: identical values)
    bn(x, bn, relu, residual)     BatchNorm2d in the module's current mode (+ residual) (+ ReLU)
    maxpool(x)                    3x3 / stride 2 / pad 1
`KernelOps` runs them on the tcgen05 GEMM / implicit-GEMM conv (csrc/gemm_tc.cu) and the NHWC streaming kernels
of csrc/maskconv.cu; BatchNorm's affine apply is torch elementwise this round.
"""
import torch
import torch.nn.functional as F

import motifs_cabi as _c
from lib import tc_ops


def bottleneck_forward(blk, x, ops):
    """torchvision Bottleneck (stride on the 3x3 conv, as the reference's lib/resnet.py:9-46)."""
    out = ops.bn(ops.conv1x1(x, blk.conv1), blk.bn1, True)
    out = ops.bn(ops.conv3x3(out, blk.conv2), blk.bn2, True)
    out = ops.conv1x1(out, blk.conv3)
    res = x if blk.downsample is None else ops.bn(ops.conv1x1(x, blk.downsample[0]), blk.downsample[1], False)
    return ops.bn(out, blk.bn3, True, residual=res)


def resnet_c4_forward(model, x, ops):
    """x [B,3,S,S] NCHW fp32 -> C4 feature map, NHWC fp32 [B,S/16,S/16,1024]."""
    y = ops.stem(x, model.conv1)
    y = ops.bn(y, model.bn1, True)
    y = ops.maxpool(y)
    for layer in (model.layer1, model.layer2, model.layer3):
        for blk in layer:
            y = bottleneck_forward(blk, y, ops)
    return y


class KernelOps(object):
    """The five operations on this library's kernels (CUDA tensors only; forward, no autograd)."""

    def stem(self, x, conv):
        _c.require_cuda(x)
        assert conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3) and conv.bias is None
        B, Cin, H, W = x.shape
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        cols = F.unfold(x, 7, padding=3, stride=2)                                  # [B, Cin*49, Ho*Wo], k = (c, ky, kx)
        a = tc_ops.split_rows(cols.transpose(1, 2).reshape(B * Ho * Wo, Cin * 49))
        w = tc_ops._cached(conv.weight, "stem7", lambda t: tc_ops.split_rows(t.reshape(t.size(0), -1)))
        return tc_ops.gemm(a, w).view(B, Ho, Wo, conv.out_channels)

    def conv1x1(self, x, conv):
        assert conv.kernel_size == (1, 1) and conv.padding == (0, 0)
        if conv.stride == (2, 2):
            x = x[:, ::2, ::2, :].contiguous()
        else:
            assert conv.stride == (1, 1)
        B, H, W, C = x.shape
        w = tc_ops._cached(conv.weight, "conv1x1", lambda t: tc_ops.split_rows(t.reshape(t.size(0), -1)))
        y = tc_ops.gemm(tc_ops.split_rows(x.reshape(-1, C)), w, bias=conv.bias.detach() if conv.bias is not None else None)
        return y.view(B, H, W, conv.out_channels)

    def conv3x3(self, x, conv):
        assert conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2))
        B, H, W, C = x.shape
        assert C % 64 == 0, "the implicit-GEMM conv reads 64-channel K blocks"
        xs = tc_ops.split_rows(x.reshape(-1, C))
        y, _ = tc_ops.conv3x3_relu((xs.hi.view(B, H, W, C), xs.lo.view(B, H, W, C)), B, H, W, C, conv,
                                   want_f32=True, want_split=False, relu=False)
        if conv.stride == (2, 2):
            y = y[:, ::2, ::2, :].contiguous()
        return y

    def bn(self, x, bn, relu, residual=None):
        from lib import mask_conv
        C = x.size(-1)
        if bn.training:
            mean, invstd = mask_conv._bn_stats(x.reshape(-1, C), bn.eps, bn.momentum, bn.running_mean, bn.running_var)
            with torch.no_grad():
                bn.num_batches_tracked += 1
        else:
            mean, invstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
        y = (x - mean) * (invstd * bn.weight.detach()) + bn.bias.detach()
        if residual is not None:
            y = y + residual
        return torch.relu_(y) if relu else y

    def maxpool(self, x):
        _c.require_cuda(x)
        B, H, W, C = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        dev = x.device
        y = torch.empty(B, Ho, Wo, C, dtype=torch.float32, device=dev)
        arg = torch.empty(B, Ho, Wo, C, dtype=torch.uint8, device=dev)
        zero = torch.zeros(C, dtype=torch.float32, device=dev)
        one = torch.ones(C, dtype=torch.float32, device=dev)
        x = x.contiguous()
        with torch.cuda.device(dev):       # BN stage of the fused kernel set to the identity: (v - 0) * 1 * 1 + 0 == v
            _c.check(_c.load().mb200_bn_pool3s2_nhwc(_c.ptr(x), _c.ptr(zero), _c.ptr(one), _c.ptr(one), _c.ptr(zero),
                                                     B, H, W, C, _c.ptr(y), _c.ptr(arg), _c.cur_stream()),
                     "mb200_bn_pool3s2_nhwc")
        return y
