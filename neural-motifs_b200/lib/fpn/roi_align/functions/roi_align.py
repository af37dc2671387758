"""RoIAlign operator over the sm_100a kernels (csrc/roi_align.cu).

Same surface as the reference's lib/fpn/roi_align/functions/roi_align.py:9-74:
`RoIAlignFunction(aligned_height, aligned_width, spatial_scale)(features, rois)` with
features [B,C,H,W] fp32 CUDA, rois [N,5] = (image idx, x1, y1, x2, y2) in image pixels;
returns [N,C,ah,aw]; gradient flows to `features` only (:74).  CPU tensors raise, as in
the reference (:45-46).
"""
import torch
from torch.autograd import Function

import motifs_cabi as _c


_SCALE_CACHE = {}


def normalize_rois(rois, feat_h, feat_w, spatial_scale):
    """roi_align.py:20-31 — corners divided by (W-1)/scale, (H-1)/scale (fp32)."""
    height = (feat_h - 1) / spatial_scale
    width = (feat_w - 1) / spatial_scale
    key = (rois.device, rois.dtype, feat_h, feat_w, spatial_scale)
    scale = _SCALE_CACHE.get(key)
    if scale is None:       # built once: `new_tensor(list)` is a pageable H2D copy, which stalls the host behind the stream
        scale = _SCALE_CACHE[key] = rois.new_tensor([1.0, width, height, width, height])
    return (rois / scale).contiguous()


class _RoIAlign(Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        _c.require_cuda(features, rois)
        if rois.dim() != 2 or rois.size(1) != 5:
            raise AssertionError("rois must be [N,5]")  # roi_align_cuda.c:19-22 returns 0 -> assert res == 1
        features = features.contiguous().float()
        rois = rois.contiguous().float()
        B, C, H, W = features.shape
        rois_n = normalize_rois(rois, H, W, spatial_scale)
        out = torch.empty(rois.size(0), C, aligned_height, aligned_width, device=features.device, dtype=torch.float32)
        lib = _c.load()
        with torch.cuda.device(features.device):
            rc = lib.ROIAlignForwardLaucher(_c.ptr(features), _c.ptr(rois_n), rois.size(0), B, H, W,
                                            aligned_height, aligned_width, C, 0.0, _c.ptr(out), _c.cur_stream())
        _c.check(rc, "ROIAlignForwardLaucher")
        ctx.save_for_backward(rois_n)
        ctx.feature_size = (B, C, H, W)
        ctx.crop = (aligned_height, aligned_width)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (rois_n,) = ctx.saved_tensors
        B, C, H, W = ctx.feature_size
        ah, aw = ctx.crop
        grad_output = grad_output.contiguous().float()
        grad_input = torch.zeros(B, C, H, W, device=grad_output.device, dtype=torch.float32)
        lib = _c.load()
        with torch.cuda.device(grad_output.device):
            rc = lib.ROIAlignBackwardLaucher(_c.ptr(grad_output), _c.ptr(rois_n), rois_n.size(0), B, H, W,
                                             ah, aw, C, _c.ptr(grad_input), _c.cur_stream())
        _c.check(rc, "ROIAlignBackwardLaucher")
        return grad_input, None, None, None, None


class RoIAlignFunction(object):
    """Legacy-style callable: constructed with the pooling geometry, then applied."""

    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIAlign.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)


def roi_align_nhwc(features_nhwc, rois, aligned_height, aligned_width, spatial_scale):
    """Pipeline variant (no autograd): features [B,H,W,C] -> pooled [N, ah*aw, C]."""
    _c.require_cuda(features_nhwc, rois)
    B, H, W, C = features_nhwc.shape
    rois_n = normalize_rois(rois.contiguous().float(), H, W, spatial_scale)
    out = torch.empty(rois.size(0), aligned_height * aligned_width, C, device=features_nhwc.device,
                      dtype=torch.float32)
    lib = _c.load()
    with torch.cuda.device(features_nhwc.device):
        rc = lib.mb200_roi_align_forward_nhwc(_c.ptr(features_nhwc), _c.ptr(rois_n), rois.size(0), B, H, W,
                                              aligned_height, aligned_width, C, 0.0, _c.ptr(out), _c.cur_stream())
    _c.check(rc, "mb200_roi_align_forward_nhwc")
    return out


def roi_align_from_nhwc(features_nhwc, rois, aligned_height, aligned_width, spatial_scale):
    """Pipeline variant (no autograd): features [B,H,W,C] (C % 4 == 0) -> the reference's
    [N,C,ah,aw] layout, so fc6's flatten order (c, y, x) is unchanged."""
    _c.require_cuda(features_nhwc, rois)
    B, H, W, C = features_nhwc.shape
    rois_n = normalize_rois(rois.detach().contiguous().float(), H, W, spatial_scale)
    out = torch.empty(rois.size(0), C, aligned_height, aligned_width, device=features_nhwc.device,
                      dtype=torch.float32)
    lib = _c.load()
    with torch.cuda.device(features_nhwc.device):
        rc = lib.mb200_roi_align_forward_nhwc_to_nchw(_c.ptr(features_nhwc), _c.ptr(rois_n), rois.size(0), B, H, W,
                                                      aligned_height, aligned_width, C, 0.0, _c.ptr(out),
                                                      _c.cur_stream())
    _c.check(rc, "mb200_roi_align_forward_nhwc_to_nchw")
    return out
