"""Train-time relation sampling for SGCls / PredCls — lib/fpn/proposal_assignments/
proposal_assignments_gtbox.py:9-87 of the reference: every GT relation is a foreground triple
(subsampled to RELS_PER_IMG*REL_FG_FRACTION*num_im), every other same-image ordered pair a
background candidate (subsampled so the total is <= RELS_PER_IMG*num_im); result sorted by
(image, subject, object). Index math runs on the device; the subsampling indices come from an
injectable numpy RNG exactly like the reference's `random_choose` (np.random.choice)."""
import numpy as np
import torch

from config import RELS_PER_IMG, REL_FG_FRACTION
from lib.pytorch_misc import random_choose


def proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, image_offset, fg_thresh=0.5, rng=np.random,
                               num_im=None):
    """rois [N,5]; gt_classes [N,2] (global image idx, class); gt_rels [R,4] (global image idx,
    subj, obj, predicate) with box indices local to the image. Returns (rois, labels [N],
    rel_labels [n,4] = (local image idx, subj row, obj row, predicate))."""
    im_inds = rois[:, 0].long()
    if num_im is None:                       # callers that hold the host copy of the image indices pass it in
        num_im = int(im_inds[-1]) + 1
    n = im_inds.size(0)
    fg_rels = gt_rels.clone()
    fg_rels[:, 0] -= image_offset
    # row offset of each image's first box: boxes are grouped by image, so it is a prefix count
    counts = torch.bincount(im_inds, minlength=num_im)
    first_row = torch.cumsum(counts, 0) - counts
    fg_rels[:, 1:3] += first_row[fg_rels[:, 0]][:, None]

    is_cand = im_inds[:, None] == im_inds[None]
    is_cand.fill_diagonal_(False)
    is_cand.view(-1)[fg_rels[:, 1] * n + fg_rels[:, 2]] = False
    is_bgcand = is_cand.nonzero()

    num_fg = min(fg_rels.size(0), int(RELS_PER_IMG * REL_FG_FRACTION * num_im))
    if num_fg < fg_rels.size(0):
        fg_rels = random_choose(fg_rels, num_fg, rng)
    num_bg = min(is_bgcand.size(0), int(RELS_PER_IMG * num_im) - num_fg)
    if num_bg > 0:
        bg_rels = torch.cat((im_inds[is_bgcand[:, 0]][:, None], is_bgcand,
                             torch.zeros(is_bgcand.size(0), 1, dtype=torch.long, device=rois.device)), 1)
        if num_bg < is_bgcand.size(0):
            bg_rels = random_choose(bg_rels, num_bg, rng)
        rel_labels = torch.cat((fg_rels, bg_rels), 0)
    else:
        rel_labels = fg_rels
    G = gt_boxes.size(0)
    _, perm = torch.sort(rel_labels[:, 0] * (G ** 2) + rel_labels[:, 1] * G + rel_labels[:, 2])
    rel_labels = rel_labels[perm].contiguous()
    labels = gt_classes[:, 1].contiguous()
    return rois, labels, rel_labels
