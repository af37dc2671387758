"""Union-box mask rasteriser. Surface of the reference's Cython module
lib/draw_rectangles/draw_rectangles.pyx:12 (`draw_union_boxes(pairs[N,8] np, pooling_size)` ->
[N,2,P,P] float32 np) plus a device-resident variant that removes the reference's mid-forward
D2H -> CPU -> H2D round trip (lib/get_union_boxes.py:47-50)."""
import numpy as np
import torch

import motifs_cabi as _c


def draw_union_boxes_cuda(pair_boxes, pooling_size, offset=0.0):
    """pair_boxes [N,8] fp32 CUDA -> [N,2,P,P] fp32 CUDA (mask - offset)."""
    _c.require_cuda(pair_boxes)
    pair_boxes = pair_boxes.contiguous().float()
    n = pair_boxes.size(0)
    out = torch.empty(n, 2, pooling_size, pooling_size, device=pair_boxes.device, dtype=torch.float32)
    lib = _c.load()
    with torch.cuda.device(pair_boxes.device):
        rc = lib.mb200_draw_union_boxes(_c.ptr(pair_boxes), n, int(pooling_size), float(offset), _c.ptr(out),
                                        _c.cur_stream())
    _c.check(rc, "mb200_draw_union_boxes")
    return out


def draw_union_boxes(bbox_pairs, pooling_size, padding=0):
    assert padding == 0, "Padding>0 not supported yet"
    pairs = torch.from_numpy(np.ascontiguousarray(bbox_pairs, dtype=np.float32)).cuda()
    return draw_union_boxes_cuda(pairs, int(pooling_size)).cpu().numpy()
