"""float64 pairwise IoU / intersection-over-query-area with numpy in and out — the surface of
the reference's Cython module lib/fpn/box_intersections_cpu/bbox.pyx:15,64 — computed by the
sm_100a kernel `mb200_bbox_overlaps_f64` (csrc/boxes.cu), bit-identical float64 results."""
import numpy as np
import torch

import motifs_cabi as _c


def _run(boxes, query_boxes, mode):
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).cuda()
    q = torch.from_numpy(np.ascontiguousarray(query_boxes, dtype=np.float64)).cuda()
    out = torch.zeros(b.size(0), q.size(0), dtype=torch.float64, device=b.device)
    lib = _c.load()
    rc = lib.mb200_bbox_overlaps_f64(_c.ptr(b), b.size(0), _c.ptr(q), q.size(0), mode, _c.ptr(out), _c.cur_stream())
    _c.check(rc, "mb200_bbox_overlaps_f64")
    return out.cpu().numpy()


def bbox_overlaps(boxes, query_boxes):
    return _run(boxes, query_boxes, 0)


def bbox_intersections(boxes, query_boxes):
    return _run(boxes, query_boxes, 1)


def bbox_overlaps_cuda(boxes, query_boxes, mode=0):
    """Device-resident variant: float64 CUDA tensors in, float64 CUDA tensor out."""
    _c.require_cuda(boxes, query_boxes)
    b = boxes.contiguous().double()
    q = query_boxes.contiguous().double()
    out = torch.zeros(b.size(0), q.size(0), dtype=torch.float64, device=b.device)
    lib = _c.load()
    with torch.cuda.device(b.device):
        rc = lib.mb200_bbox_overlaps_f64(_c.ptr(b), b.size(0), _c.ptr(q), q.size(0), mode, _c.ptr(out), _c.cur_stream())
    _c.check(rc, "mb200_bbox_overlaps_f64")
    return out
