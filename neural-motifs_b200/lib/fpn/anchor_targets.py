"""RPN anchor-target assignment — lib/fpn/anchor_targets.py:16-105 of the reference (numpy in
the DataLoader collate there). Here the 27 380 x G float64 IoU, the per-anchor max / first arg-max,
the per-GT maxima and the labelling run as two warp-per-anchor kernels (`mb200_anchor_targets`,
csrc/boxes.cu: shuffle reductions, 64-bit atomicMax for the column maxima, the IoU matrix is never
written); results are bit-identical to bbox.pyx + numpy. Only the fg/bg subsampling (npr.choice,
injectable, consumed in the reference's order) stays on the host."""
import numpy as np
import numpy.random as npr
import torch

from config import IM_SCALE, RPN_NEGATIVE_OVERLAP, RPN_POSITIVE_OVERLAP, RPN_BATCHSIZE, RPN_FG_FRACTION, \
    ANCHOR_SIZE, ANCHOR_SCALES, ANCHOR_RATIOS
import motifs_cabi as _c
from lib.fpn.generate_anchors import generate_anchors

_ANCHORS = None


def _anchors():
    global _ANCHORS
    if _ANCHORS is None:
        _ANCHORS = generate_anchors(base_size=ANCHOR_SIZE, feat_stride=16, anchor_scales=ANCHOR_SCALES,
                                    anchor_ratios=ANCHOR_RATIOS)
    return _ANCHORS


def anchor_labels_device(anchors, gt_boxes, neg_thr=RPN_NEGATIVE_OVERLAP, pos_thr=RPN_POSITIVE_OVERLAP):
    """anchors [N,4], gt_boxes [G,4] float64 CUDA -> (labels int64 [N] in {-1,0,1} before subsampling, first arg-max
    int32 [N], max overlap float64 [N]), all on the device (anchor_targets.py:50-67)."""
    _c.require_cuda(anchors, gt_boxes)
    a = anchors.contiguous().double(); g = gt_boxes.contiguous().double()
    N, G = a.size(0), g.size(0)
    dev = a.device
    labels = torch.empty(N, dtype=torch.long, device=dev)
    arg = torch.empty(N, dtype=torch.int32, device=dev)
    mx = torch.empty(N, dtype=torch.float64, device=dev)
    ws = torch.empty(max(G, 1), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _c.load().mb200_anchor_targets(_c.ptr(a), N, _c.ptr(g), G, float(neg_thr), float(pos_thr), _c.ptr(ws),
                                            _c.ptr(mx), _c.ptr(arg), _c.ptr(labels), _c.cur_stream())
    _c.check(rc, "mb200_anchor_targets")
    return labels, arg, mx


def anchor_target_layer(gt_boxes, im_size, allowed_border=0, rng=npr):
    if max(im_size) != IM_SCALE:
        raise ValueError("im size is {}".format(im_size))
    h, w = im_size
    ans_np = _anchors()
    flat = ans_np.reshape((-1, 4))
    inds_inside = np.where((flat[:, 0] >= -allowed_border) & (flat[:, 1] >= -allowed_border) &
                           (flat[:, 2] < w + allowed_border) & (flat[:, 3] < h + allowed_border))[0]
    good = flat[inds_inside]
    if good.size == 0:
        raise ValueError("There were no good anchors for an image of size {} with boxes {}".format(im_size, gt_boxes))
    gt_boxes = np.asarray(gt_boxes)
    labels, anchor_to_gtbox, _ = anchor_labels_device(torch.from_numpy(good).cuda(),
                                                      torch.from_numpy(gt_boxes.astype(np.float64)).cuda())
    labels = labels.cpu().numpy()
    anchor_to_gtbox = anchor_to_gtbox.cpu().numpy().astype(np.int64)

    num_fg = int(RPN_FG_FRACTION * RPN_BATCHSIZE)
    fg_inds = np.where(labels == 1)[0]
    if len(fg_inds) > num_fg:
        labels[rng.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
    num_bg = RPN_BATCHSIZE - np.sum(labels == 1)
    bg_inds = np.where(labels == 0)[0]
    if len(bg_inds) > num_bg:
        labels[rng.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1

    labels_unmap = (-1) * np.ones(flat.shape[0], dtype=np.int64)
    labels_unmap[inds_inside] = labels
    anchor_inds = np.column_stack(np.where(labels_unmap.reshape(ans_np.shape[:-1]) >= 0))
    sel = np.where(labels >= 0)[0]
    return good[sel], anchor_inds, gt_boxes[anchor_to_gtbox[sel]], labels[sel]
