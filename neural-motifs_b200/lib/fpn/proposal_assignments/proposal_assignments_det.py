"""Detector train-time roi sampling — lib/fpn/proposal_assignments/proposal_assignments_det.py:12-117 of
the reference: proposals + GT boxes per image, IoU arg-max against the image's GT boxes (fp32 device
kernel), <= 64 foreground (IoU >= 0.5) and the rest background (IoU in [0, 0.5)) up to 256 rois per
image; background labels clamped to 0. Sampling RNG injectable (npr.choice, reference order)."""
import numpy as np
import numpy.random as npr
import torch

from config import BG_THRESH_HI, BG_THRESH_LO, FG_FRACTION, ROIS_PER_IMG
from lib.fpn.box_utils import bbox_overlaps
from lib.pytorch_misc import to_device_async


def _sel_inds(max_overlaps, fg_thresh=0.5, fg_rois_per_image=128, rois_per_image=256, rng=npr):
    fg_inds = np.where(max_overlaps >= fg_thresh)[0]
    n_fg = min(fg_rois_per_image, fg_inds.shape[0])
    if fg_inds.size > 0:
        fg_inds = rng.choice(fg_inds, size=n_fg, replace=False)
    bg_inds = np.where((max_overlaps < BG_THRESH_HI) & (max_overlaps >= BG_THRESH_LO))[0]
    n_bg = min(rois_per_image - n_fg, bg_inds.size)
    if bg_inds.size > 0:
        bg_inds = rng.choice(bg_inds, size=n_bg, replace=False)
    return np.append(fg_inds, bg_inds), n_fg


def proposal_assignments_det(rpn_rois, gt_boxes, gt_classes, image_offset, fg_thresh=0.5, rng=npr):
    fg_rois_per_image = int(np.round(ROIS_PER_IMG * FG_FRACTION))
    dev = rpn_rois.device
    gt_img_inds = gt_classes[:, 0] - image_offset
    all_boxes = torch.cat([rpn_rois[:, 1:], gt_boxes], 0)
    ims_per_box = torch.cat([rpn_rois[:, 0].long(), gt_img_inds], 0)
    im_sorted, idx = torch.sort(ims_per_box, dim=0, stable=True)
    all_boxes = all_boxes[idx]
    im_np = im_sorted.cpu().numpy()
    gt_np = gt_img_inds.cpu().numpy()
    num_images = int(im_np[-1]) + 1
    labels, rois, bbox_targets = [], [], []
    for im in range(num_images):
        g = np.where(gt_np == im)[0]
        if g.size == 0:
            continue
        g_start, g_end = int(g[0]), int(g[-1]) + 1
        t = np.where(im_np == im)[0]
        t_start, t_end = int(t[0]), int(t[-1]) + 1
        ious = bbox_overlaps(all_boxes[t_start:t_end].contiguous(), gt_boxes[g_start:g_end].contiguous())
        max_overlaps = ious.max(1)[0]
        gt_assignment = ious.argmax(1) + g_start       # first maximum, as numpy (torch.max's index is unspecified on ties)
        keep_np, num_fg = _sel_inds(max_overlaps.cpu().numpy(), fg_thresh, fg_rois_per_image, ROIS_PER_IMG, rng)
        if keep_np.size == 0:
            continue
        keep = to_device_async(keep_np, dev, torch.long)
        labels_ = gt_classes[:, 1][gt_assignment[keep]].clone()
        if num_fg < labels_.size(0):
            labels_[num_fg:] = 0
        rois.append(torch.cat((im_sorted[t_start:t_end, None][keep].float(), all_boxes[t_start:t_end][keep]), 1))
        labels.append(labels_)
        bbox_targets.append(gt_boxes[gt_assignment[keep]])
    return torch.cat(rois, 0), torch.cat(labels, 0), torch.cat(bbox_targets, 0)
