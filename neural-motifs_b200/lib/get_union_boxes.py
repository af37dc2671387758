"""Union-box features — same surface as the reference's lib/get_union_boxes.py:15-93:
`UnionBoxesAndFeats(pooling_size, stride, dim)(fmap, rois, union_inds)` = RoIAlign over the union
roi of each (subject, object) pair + a small conv net over the pair's two rasterised box masks.

The reference goes D2H -> Cython rasteriser on the CPU -> H2D in the middle of forward (:47-50);
here union rois, pair boxes and the [N,2,27,27] masks are produced on the device (csrc/boxes.cu)."""
import torch
from torch import nn
from torch.nn.modules.module import Module

import motifs_cabi as _c
from config import BATCHNORM_MOMENTUM
from lib import mask_conv
from lib.draw_rectangles.draw_rectangles import draw_union_boxes_cuda
from lib.fpn.roi_align.functions.roi_align import RoIAlignFunction, roi_align_from_nhwc


import os
_MASKCONV_CHANNELS_LAST = os.environ.get("MOTIFS_MASKCONV_CL", "0") == "1"
_MASKCONV_IMPL = os.environ.get("MOTIFS_MASKCONV", "own")   # "own": csrc/maskconv.cu + tcgen05 GEMM; "cudnn": torch modules
_CUDNN_BENCHMARK = os.environ.get("MOTIFS_CUDNN_BENCHMARK", "1") == "1"   # let cuDNN pick its fastest fp32 algorithm


def union_rois_and_pairs(rois, union_inds):
    """rois [N,5], union_inds [R,2] int64 -> (union rois [R,5], pair boxes [R,8]) in one kernel
    (get_union_boxes.py:82-87 and the gather of :47)."""
    _c.require_cuda(rois, union_inds)
    rois = rois.detach().contiguous().float()
    union_inds = union_inds.contiguous().long()
    R = union_inds.size(0)
    u = torch.empty(R, 5, device=rois.device, dtype=torch.float32)
    pb = torch.empty(R, 8, device=rois.device, dtype=torch.float32)
    with torch.cuda.device(rois.device):
        _c.check(_c.load().mb200_union_rois(_c.ptr(rois), _c.ptr(union_inds), R, _c.ptr(u), _c.ptr(pb), _c.cur_stream()),
                 "mb200_union_rois")
    return u, pb


def union_boxes(fmap, rois, union_inds, pooling_size=14, stride=16):
    """get_union_boxes.py:72-93 (gradients reach fmap through RoIAlign backward)."""
    assert union_inds.size(1) == 2
    u, _ = union_rois_and_pairs(rois, union_inds)
    return RoIAlignFunction(pooling_size, pooling_size, spatial_scale=1 / stride)(fmap, u)


class _MaxPool3s2(torch.autograd.Function):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NCHW fp32 over csrc/pool.cu."""

    @staticmethod
    def forward(ctx, x):
        _c.require_cuda(x)
        x = x.contiguous()
        N, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty(N, C, Ho, Wo, device=x.device, dtype=torch.float32)
        arg = torch.empty(N, C, Ho, Wo, device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device):
            _c.check(_c.load().mb200_maxpool3s2_forward(_c.ptr(x), N * C, H, W, _c.ptr(y), _c.ptr(arg), _c.cur_stream()),
                     "mb200_maxpool3s2_forward")
        ctx.save_for_backward(arg)
        ctx.hw = (H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        (arg,) = ctx.saved_tensors
        H, W = ctx.hw
        N, C = arg.shape[:2]
        gy = gy.contiguous()
        gx = torch.empty(N, C, H, W, device=gy.device, dtype=torch.float32)
        with torch.cuda.device(gy.device):
            _c.check(_c.load().mb200_maxpool3s2_backward(_c.ptr(gy), _c.ptr(arg), N * C, H, W, _c.ptr(gx), _c.cur_stream()),
                     "mb200_maxpool3s2_backward")
        return gx


class UnionBoxesAndFeats(Module):
    def __init__(self, pooling_size=7, stride=16, dim=256, concat=False, use_feats=True):
        super().__init__()
        self.pooling_size = pooling_size
        self.stride = stride
        self.dim = dim
        self.use_feats = use_feats
        self.conv = nn.Sequential(
            nn.Conv2d(2, dim // 2, kernel_size=7, stride=2, padding=3, bias=True),
            nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim // 2, momentum=BATCHNORM_MOMENTUM),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(dim // 2, dim, kernel_size=3, stride=1, padding=1, bias=True),
            nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim, momentum=BATCHNORM_MOMENTUM),
        )
        self.concat = concat

    def forward(self, fmap, rois, union_inds, fmap_nhwc=None):
        """fmap [B,C,H,W]; when the caller also holds the NHWC copy the backbone produced
        (`fmap_nhwc`, no gradient needed) the pooled features come from the channel-vectorised kernel."""
        u, pair_boxes = union_rois_and_pairs(rois, union_inds)
        if fmap_nhwc is not None and not fmap.requires_grad:
            union_pools = roi_align_from_nhwc(fmap_nhwc, u, self.pooling_size, self.pooling_size, 1 / self.stride)
        else:
            union_pools = RoIAlignFunction(self.pooling_size, self.pooling_size, spatial_scale=1 / self.stride)(fmap, u)
        if not self.use_feats:
            return union_pools.detach()
        rects = draw_union_boxes_cuda(pair_boxes, self.pooling_size * 4 - 1, offset=0.5)
        if _MASKCONV_IMPL == "own" and mask_conv.supported(self.conv):
            # conv7x7/s2 + ReLU + BN + pool + conv3x3 + ReLU + BN (+ the residual add) on this library's kernels
            if self.concat:
                return torch.cat((union_pools, mask_conv.mask_conv_net(self.conv, rects)), 1)
            return mask_conv.mask_conv_net(self.conv, rects, addend=union_pools)
        # MOTIFS_MASKCONV=cudnn: the torch modules (cuDNN). The mask conv net (7x7 s2 + 3x3) on cuDNN; TF32 is
        # switched off so it stays inside the fp32 parity bar (TF32 alone costs ~1e-3 here).
        if _MASKCONV_CHANNELS_LAST:
            # NHWC kernels for the cuDNN conv / BN / pool of this branch (layout only; same arithmetic)
            if not getattr(self, "_cl_done", False):
                self.conv.to(memory_format=torch.channels_last)
                self._cl_done = True
            rects = rects.contiguous(memory_format=torch.channels_last)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=_CUDNN_BENCHMARK):
            conv_out = rects
            for m in self.conv:
                if isinstance(m, nn.MaxPool2d) and m.kernel_size == 3 and m.stride == 2 and m.padding == 1 \
                        and not _MASKCONV_CHANNELS_LAST:
                    conv_out = _MaxPool3s2.apply(conv_out)
                else:
                    conv_out = m(conv_out)
        if self.concat:
            return torch.cat((union_pools, conv_out), 1)
        return union_pools + conv_out
