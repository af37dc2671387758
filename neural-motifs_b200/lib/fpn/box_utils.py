"""Box math, same names and conventions as the reference's lib/fpn/box_utils.py (the +1 pixel
convention of center_size / point_form, :51-82).  CUDA tensors go through the sm_100a kernels
in csrc/boxes.cu; numpy arrays go through the float64 IoU kernel (bbox.pyx semantics)."""
import numpy as np
import torch
from torch.nn import functional as F

import motifs_cabi as _c
from lib.fpn.box_intersections_cpu.bbox import bbox_overlaps as bbox_overlaps_np
from lib.fpn.box_intersections_cpu.bbox import bbox_intersections as bbox_intersections_np


def center_size(boxes):
    """(x1,y1,x2,y2) -> (cx,cy,w,h), box_utils.py:51-63."""
    wh = boxes[:, 2:] - boxes[:, :2] + 1.0
    if isinstance(boxes, np.ndarray):
        return np.column_stack((boxes[:, :2] + 0.5 * wh, wh))
    return torch.cat((boxes[:, :2] + 0.5 * wh, wh), 1)


def point_form(boxes):
    """(cx,cy,w,h) -> (x1,y1,x2,y2), box_utils.py:66-79."""
    if isinstance(boxes, np.ndarray):
        return np.column_stack((boxes[:, :2] - 0.5 * boxes[:, 2:], boxes[:, :2] + 0.5 * (boxes[:, 2:] - 2.0)))
    return torch.cat((boxes[:, :2] - 0.5 * boxes[:, 2:], boxes[:, :2] + 0.5 * (boxes[:, 2:] - 2.0)), 1)


def bbox_preds(boxes, deltas):
    """Delta decode, box_utils.py:28-48: one fused kernel instead of ~10 elementwise launches.
    boxes [N,4], deltas [N,4] (same row count)."""
    if boxes.size(0) == 0:
        return boxes
    if not boxes.is_cuda or boxes.requires_grad or deltas.requires_grad:
        pc = center_size(boxes)
        xys = pc[:, :2] + pc[:, 2:] * deltas[:, :2]
        whs = torch.exp(deltas[:, 2:]) * pc[:, 2:]
        return point_form(torch.cat((xys, whs), 1))
    return bbox_preds_fused(boxes, deltas, 1)


def bbox_preds_fused(boxes, deltas, rows_per_box, im_hw=None, im_idx=None):
    """boxes [N,4]; deltas [N*K,4]; optional clamp to each row's image (h,w)."""
    _c.require_cuda(boxes, deltas)
    boxes = boxes.contiguous().float()
    deltas = deltas.contiguous().float()
    out = torch.empty_like(deltas)
    lib = _c.load()
    with torch.cuda.device(boxes.device):
        rc = lib.mb200_bbox_preds(_c.ptr(boxes), _c.ptr(deltas), deltas.size(0), int(rows_per_box),
                                  _c.ptr(im_hw), _c.ptr(im_idx), _c.ptr(out), _c.cur_stream())
    _c.check(rc, "mb200_bbox_preds")
    return out


def bbox_intersections(box_a, box_b):
    """box_utils.py:85-106."""
    if isinstance(box_a, np.ndarray):
        assert isinstance(box_b, np.ndarray)
        return bbox_intersections_np(box_a, box_b)
    max_xy = torch.min(box_a[:, None, 2:], box_b[None, :, 2:])
    min_xy = torch.max(box_a[:, None, :2], box_b[None, :, :2])
    inter = torch.clamp((max_xy - min_xy + 1.0), min=0)
    return inter[:, :, 0] * inter[:, :, 1]


def bbox_overlaps(box_a, box_b):
    """Pairwise IoU [A,B], box_utils.py:109-131."""
    if isinstance(box_a, np.ndarray):
        assert isinstance(box_b, np.ndarray)
        return bbox_overlaps_np(box_a, box_b)
    _c.require_cuda(box_a, box_b)
    a = box_a.detach().contiguous().float()
    b = box_b.detach().contiguous().float()
    out = torch.empty(a.size(0), b.size(0), device=a.device, dtype=torch.float32)
    lib = _c.load()
    with torch.cuda.device(a.device):
        rc = lib.mb200_bbox_overlaps_f32(_c.ptr(a), a.size(0), _c.ptr(b), b.size(0), _c.ptr(out), _c.cur_stream())
    _c.check(rc, "mb200_bbox_overlaps_f32")
    return out


def nms_overlaps(boxes):
    """Per-class pairwise IoU, boxes [N,nc,4] -> [N,N,nc], box_utils.py:134-154."""
    assert boxes.dim() == 3
    max_xy = torch.min(boxes[:, None, :, 2:], boxes[None, :, :, 2:])
    min_xy = torch.max(boxes[:, None, :, :2], boxes[None, :, :, :2])
    inter = torch.clamp((max_xy - min_xy + 1.0), min=0)
    inters = inter[..., 0] * inter[..., 1]
    areas = (boxes[..., 2] - boxes[..., 0] + 1.0) * (boxes[..., 3] - boxes[..., 1] + 1.0)
    union = -inters + areas[None] + areas[:, None]
    return inters / union


def bbox_loss(prior_boxes, deltas, gt_boxes, eps=1e-4, scale_before=1):
    """Smooth-L1 box regression loss, box_utils.py:8-25."""
    prior_centers = center_size(prior_boxes)
    gt_centers = center_size(gt_boxes)
    center_targets = (gt_centers[:, :2] - prior_centers[:, :2]) / prior_centers[:, 2:]
    size_targets = torch.log(gt_centers[:, 2:]) - torch.log(prior_centers[:, 2:])
    all_targets = torch.cat((center_targets, size_targets), 1)
    return F.smooth_l1_loss(deltas, all_targets, reduction='sum') / (eps + prior_centers.size(0))
