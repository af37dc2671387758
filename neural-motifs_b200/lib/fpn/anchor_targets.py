"""RPN anchor-target assignment — lib/fpn/anchor_targets.py:16-105 of the reference (numpy in
the DataLoader collate there). Here the 27 380 x G float64 IoU runs on the device
(`mb200_bbox_overlaps_f64`, bit-identical to bbox.pyx) and the arg-max / labelling is done with
torch on the device; only the fg/bg subsampling (npr.choice, injectable) stays on the host."""
import numpy as np
import numpy.random as npr
import torch

from config import IM_SCALE, RPN_NEGATIVE_OVERLAP, RPN_POSITIVE_OVERLAP, RPN_BATCHSIZE, RPN_FG_FRACTION, \
    ANCHOR_SIZE, ANCHOR_SCALES, ANCHOR_RATIOS
from lib.fpn.box_intersections_cpu.bbox import bbox_overlaps_cuda
from lib.fpn.generate_anchors import generate_anchors

_ANCHORS = None


def _anchors():
    global _ANCHORS
    if _ANCHORS is None:
        _ANCHORS = generate_anchors(base_size=ANCHOR_SIZE, feat_stride=16, anchor_scales=ANCHOR_SCALES,
                                    anchor_ratios=ANCHOR_RATIOS)
    return _ANCHORS


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
    ov = bbox_overlaps_cuda(torch.from_numpy(good).cuda(), torch.from_numpy(gt_boxes.astype(np.float64)).cuda())
    max_overlaps, anchor_to_gtbox = ov.max(1)
    gt_max = ov.max(0)[0]
    is_gt_argmax = (ov == gt_max[None]).any(1)
    labels = torch.full((ov.size(0),), -1, dtype=torch.long, device=ov.device)
    labels[max_overlaps < RPN_NEGATIVE_OVERLAP] = 0
    labels[is_gt_argmax] = 1
    labels[max_overlaps >= RPN_POSITIVE_OVERLAP] = 1
    # numpy's argmax takes the FIRST maximum; torch.max on CUDA does not promise that, so re-derive it
    first_arg = (ov == max_overlaps[:, None]).to(torch.uint8).argmax(1)
    labels = labels.cpu().numpy()
    anchor_to_gtbox = first_arg.cpu().numpy()

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
