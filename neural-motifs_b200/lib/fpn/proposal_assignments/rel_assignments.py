"""SGDet train-time relation labels — lib/fpn/proposal_assignments/rel_assignments.py:15-145 of the
reference: detections are matched to GT boxes (same class, IoU >= 0.5), every GT relation spawns
foreground candidates between matching detections (sampled with probability proportional to the IoU
product), background candidates are the remaining ordered pairs of labelled, overlapping, non-identical
detections; <= 16 fg and 64 total per image; sorted by (subject, object).

IoUs come from the float64 device kernel (bit-identical to the reference's Cython); the candidate
enumeration is vectorised; the random draws use an injectable numpy RNG in the reference's call order
(`choice(p=...)` per GT relation, `choice` for the fg cap, `choice` for bg)."""
import numpy as np
import numpy.random as npr
import torch

from config import REL_FG_FRACTION
from lib.fpn.box_intersections_cpu.bbox import bbox_overlaps_cuda
from lib.pytorch_misc import to_device_async


def rel_assignments(im_inds, rpn_rois, roi_gtlabels, gt_boxes, gt_classes, gt_rels, image_offset, fg_thresh=0.5,
                    num_sample_per_gt=4, filter_non_overlap=True, rng=npr):
    fg_rels_per_image = int(np.round(REL_FG_FRACTION * 64))
    dev = rpn_rois.device
    pred_inds = im_inds.cpu().numpy()
    pred_labels = roi_gtlabels.cpu().numpy()
    gt_cls = gt_classes.cpu().numpy().copy()
    gt_rel = gt_rels.cpu().numpy().copy()
    gt_cls[:, 0] -= image_offset
    gt_rel[:, 0] -= image_offset
    num_im = int(gt_cls[:, 0].max()) + 1
    boxes64 = rpn_rois.detach().double()
    gt64 = gt_boxes.detach().double()

    out = []
    num_box_seen = 0
    for im in range(num_im):
        p_idx = np.where(pred_inds == im)[0]
        g_idx = np.where(gt_cls[:, 0] == im)[0]
        n = p_idx.shape[0]
        pl = pred_labels[p_idx]
        gcls = gt_cls[g_idx, 1]
        rels_i = gt_rel[gt_rel[:, 0] == im, 1:]
        pb = boxes64[torch.as_tensor(p_idx, device=dev)]
        ious = bbox_overlaps_cuda(pb, gt64[torch.as_tensor(g_idx, device=dev)]).cpu().numpy() if n and len(g_idx) \
            else np.zeros((n, len(g_idx)))
        self_iou = bbox_overlaps_cuda(pb, pb).cpu().numpy() if n else np.zeros((0, 0))
        is_match = (pl[:, None] == gcls[None]) & (ious >= fg_thresh)
        if filter_non_overlap:
            possible = (self_iou < 1) & (self_iou > 0)
        else:
            possible = ~np.eye(n, dtype=bool)
        possible = possible.copy()
        possible[pl == 0] = False
        possible[:, pl == 0] = False

        fg = []
        for (g_from, g_to, rel_id) in rels_i:
            fr = np.where(is_match[:, g_from])[0]
            to = np.where(is_match[:, g_to])[0]
            if fr.size == 0 or to.size == 0:
                continue
            ff, tt = np.meshgrid(fr, to, indexing='ij')          # from-major, as the nested loops
            ff, tt = ff.reshape(-1), tt.reshape(-1)
            keep = ff != tt
            ff, tt = ff[keep], tt[keep]
            if ff.size == 0:
                continue
            possible[ff, tt] = False
            p = ious[ff, g_from] * ious[tt, g_to]
            p = p / p.sum()
            for k in rng.choice(p.shape[0], p=p, size=min(p.shape[0], num_sample_per_gt), replace=False):
                fg.append((ff[k], tt[k], rel_id))
        fg = np.array(fg, dtype=np.int64).reshape(-1, 3)
        if fg.shape[0] > fg_rels_per_image:
            fg = fg[rng.choice(fg.shape[0], size=fg_rels_per_image, replace=False)]

        bg = np.column_stack(np.where(possible))
        bg = np.column_stack((bg, np.zeros(bg.shape[0], dtype=np.int64)))
        num_bg = min(64 - fg.shape[0], bg.shape[0])
        if bg.size > 0:
            bg = bg[rng.choice(bg.shape[0], size=num_bg, replace=False)]
        else:
            bg = np.zeros((0, 3), dtype=np.int64)
        if fg.size == 0 and bg.size == 0:
            bg = np.array([[0, 0, 0]], dtype=np.int64)       # "just put something here" (:126-128)
        allr = np.concatenate((fg, bg), 0)
        allr[:, 0:2] += num_box_seen
        allr = allr[np.lexsort((allr[:, 1], allr[:, 0]))]
        out.append(np.column_stack((im * np.ones(allr.shape[0], dtype=np.int64), allr)))
        num_box_seen += n
    return to_device_async(np.concatenate(out, 0), dev, torch.long)
