"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the host-side helpers of the path:
anchors, RPN anchor-target assignment, sequence packing indices, relation sampling."""
import numpy as np

from . import ops

IM_SCALE = 592
ANCHOR_SIZE = 16
ANCHOR_RATIOS = (0.23232838, 0.63365731, 1.28478321, 3.15089189)
ANCHOR_SCALES = (2.22152954, 4.12315647, 7.21692515, 12.60263013, 22.7102731)
RPN_POSITIVE_OVERLAP, RPN_NEGATIVE_OVERLAP = 0.7, 0.3
RPN_FG_FRACTION, RPN_BATCHSIZE = 0.5, 256
RELS_PER_IMG, REL_FG_FRACTION = 256, 0.25


def _whctrs(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mk(ws, hs, xc, yc):
    ws, hs = ws[:, None], hs[:, None]
    return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))


def generate_base_anchors(base_size=16, ratios=ANCHOR_RATIOS, scales=ANCHOR_SCALES):
    """lib/fpn/generate_anchors.py:62-126 — ratio enumeration WITHOUT rounding (:110-111)."""
    ratios, scales = np.array(ratios), np.array(scales)
    base = np.array([1, 1, base_size, base_size]) - 1
    w, h, xc, yc = _whctrs(base)
    ws = np.sqrt(w * h / ratios)
    hs = ws * ratios
    ratio_anchors = _mk(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):
        w, h, xc, yc = _whctrs(ratio_anchors[i])
        out.append(_mk(w * scales, h * scales, xc, yc))
    return np.vstack(out)


def generate_anchors(base_size=ANCHOR_SIZE, feat_stride=16, anchor_scales=ANCHOR_SCALES, anchor_ratios=ANCHOR_RATIOS):
    """lib/fpn/generate_anchors.py:39-52 -> [37,37,A,4] float64."""
    anchors = generate_base_anchors(base_size, anchor_ratios, anchor_scales)
    shift = np.arange(0, IM_SCALE // feat_stride) * feat_stride
    sx, sy = np.meshgrid(shift, shift)
    shifts = np.stack([sx, sy, sx, sy], -1)
    return shifts[:, :, None] + anchors[None, None]


def anchor_target_labels(gt_boxes, im_size, allowed_border=0):
    """Deterministic half of anchor_target_layer (lib/fpn/anchor_targets.py:16-71): inside-image
    filter, float64 IoU, labels before fg/bg subsampling. Returns (ans, inds_inside, labels,
    anchor_to_gtbox)."""
    h, w = im_size
    ans = generate_anchors()
    flat = ans.reshape((-1, 4))
    inds_inside = np.where((flat[:, 0] >= -allowed_border) & (flat[:, 1] >= -allowed_border) &
                           (flat[:, 2] < w + allowed_border) & (flat[:, 3] < h + allowed_border))[0]
    good = flat[inds_inside]
    overlaps = ops.bbox_overlaps_f64(good, gt_boxes)
    anchor_to_gtbox = overlaps.argmax(axis=1)
    max_overlaps = overlaps[np.arange(anchor_to_gtbox.shape[0]), anchor_to_gtbox]
    gtbox_to_anchor = overlaps.argmax(axis=0)
    gt_max_overlaps = overlaps[gtbox_to_anchor, np.arange(overlaps.shape[1])]
    gt_argmax_overlaps = np.where(overlaps == gt_max_overlaps)[0]
    labels = (-1) * np.ones(overlaps.shape[0], dtype=np.int64)
    labels[max_overlaps < RPN_NEGATIVE_OVERLAP] = 0
    labels[gt_argmax_overlaps] = 1
    labels[max_overlaps >= RPN_POSITIVE_OVERLAP] = 1
    return ans, inds_inside, labels, anchor_to_gtbox


def anchor_target_layer(gt_boxes, im_size, rng=np.random):
    """lib/fpn/anchor_targets.py:16-105, the RNG (npr.choice) injected."""
    ans, inds_inside, labels, anchor_to_gtbox = anchor_target_labels(gt_boxes, im_size)
    flat = ans.reshape((-1, 4))
    good = flat[inds_inside]
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
    anchor_inds = np.column_stack(np.where(labels_unmap.reshape(ans.shape[:-1]) >= 0))
    sel = np.where(labels >= 0)[0]
    return good[sel], anchor_inds, np.asarray(gt_boxes)[anchor_to_gtbox[sel]], labels[sel]


def enumerate_by_image(im_inds):
    """lib/pytorch_misc.py:278-287: runs of equal image index -> (image, start, end)."""
    im_inds = np.asarray(im_inds)
    out, s, cur = [], 0, int(im_inds[0])
    for i, v in enumerate(im_inds):
        if v != cur:
            out.append((cur, s, i))
            cur, s = int(v), i
    out.append((cur, s, len(im_inds)))
    return out


def transpose_packed_sequence_inds(lengths):
    """lib/pytorch_misc.py:365-384: BxT (image-major) -> TxB (time-major) gather indices and the
    per-timestep batch sizes, for lengths sorted descending."""
    new_inds, new_lens = [], []
    cum = np.cumsum([0] + list(lengths))
    ptr = len(lengths) - 1
    for i in range(lengths[0]):
        while ptr > 0 and lengths[ptr] <= i:
            ptr -= 1
        new_inds.append(cum[:ptr + 1].copy())
        cum[:ptr + 1] += 1
        new_lens.append(ptr + 1)
    return np.concatenate(new_inds, 0), new_lens


def rel_assignments(im_inds, rois, roi_gtlabels, gt_boxes, gt_classes, gt_rels, image_offset, rng, fg_thresh=0.5,
                    num_sample_per_gt=4, filter_non_overlap=True):
    """lib/fpn/proposal_assignments/rel_assignments.py:15-145 on numpy arrays, nested loops as written
    there; `rng` replaces numpy.random (same call order). Returns int64 [n,4]."""
    fg_rels_per_image = int(np.round(REL_FG_FRACTION * 64))
    gt_classes = np.array(gt_classes, copy=True); gt_rels = np.array(gt_rels, copy=True)
    gt_classes[:, 0] -= image_offset
    gt_rels[:, 0] -= image_offset
    num_im = gt_classes[:, 0].max() + 1
    rel_labels, num_box_seen = [], 0
    for im_ind in range(num_im):
        pred_ind = np.where(im_inds == im_ind)[0]
        gt_ind = np.where(gt_classes[:, 0] == im_ind)[0]
        gt_boxes_i = gt_boxes[gt_ind]
        gt_classes_i = gt_classes[gt_ind, 1]
        gt_rels_i = gt_rels[gt_rels[:, 0] == im_ind, 1:]
        pred_boxes_i = rois[pred_ind]
        labels_i = roi_gtlabels[pred_ind]
        ious = ops.bbox_overlaps_f64(pred_boxes_i, gt_boxes_i)
        is_match = (labels_i[:, None] == gt_classes_i[None]) & (ious >= fg_thresh)
        pbi = ops.bbox_overlaps_f64(pred_boxes_i, pred_boxes_i)
        if filter_non_overlap:
            poss = (pbi < 1) & (pbi > 0)
        else:
            poss = (np.ones((len(pred_ind), len(pred_ind)), dtype=np.int64) - np.eye(len(pred_ind), dtype=np.int64)) > 0
        poss = poss.copy()
        poss[labels_i == 0] = 0
        poss[:, labels_i == 0] = 0
        fg_rels = []
        for (from_gt, to_gt, rel_id) in gt_rels_i:
            cand, score = [], []
            for a in np.where(is_match[:, from_gt])[0]:
                for b in np.where(is_match[:, to_gt])[0]:
                    if a != b:
                        cand.append((a, b, rel_id))
                        score.append(ious[a, from_gt] * ious[b, to_gt])
                        poss[a, b] = 0
            if not cand:
                continue
            p = np.array(score); p = p / p.sum()
            for k in rng.choice(p.shape[0], p=p, size=min(p.shape[0], num_sample_per_gt), replace=False):
                fg_rels.append(cand[k])
        fg_rels = np.array(fg_rels, dtype=np.int64)
        if fg_rels.size > 0 and fg_rels.shape[0] > fg_rels_per_image:
            fg_rels = fg_rels[rng.choice(fg_rels.shape[0], size=fg_rels_per_image, replace=False)]
        elif fg_rels.size == 0:
            fg_rels = np.zeros((0, 3), dtype=np.int64)
        bg = np.column_stack(np.where(poss))
        bg = np.column_stack((bg, np.zeros(bg.shape[0], dtype=np.int64)))
        num_bg = min(64 - fg_rels.shape[0], bg.shape[0])
        if bg.size > 0:
            bg = bg[rng.choice(bg.shape[0], size=num_bg, replace=False)]
        else:
            bg = np.zeros((0, 3), dtype=np.int64)
        if fg_rels.size == 0 and bg.size == 0:
            bg = np.array([[0, 0, 0]], dtype=np.int64)
        allr = np.concatenate((fg_rels, bg), 0)
        allr[:, 0:2] += num_box_seen
        allr = allr[np.lexsort((allr[:, 1], allr[:, 0]))]
        rel_labels.append(np.column_stack((im_ind * np.ones(allr.shape[0], dtype=np.int64), allr)))
        num_box_seen += pred_boxes_i.shape[0]
    return np.concatenate(rel_labels, 0)


def proposal_assignments_det(rois, gt_boxes, gt_classes, image_offset, rng, fg_thresh=0.5, order=None):
    """lib/fpn/proposal_assignments/proposal_assignments_det.py:12-117 on numpy arrays (fp32 IoU as the
    torch branch of box_utils.bbox_overlaps). Returns (rois [n,5] f32, labels [n] i64, targets [n,4] f32).
    The reference orders the candidates with `torch.sort(ims_per_box, 0)` (:33), whose order among equal image
    indices is implementation-defined (torch 2.11's CPU sort is NOT stable there); this restatement and the product
    use the stable order. `order` injects a given permutation so that the fixture produced by running the reference
    (tests/golden/make_golden_host2.py stores the permutation its torch.sort returned) can be matched exactly."""
    fg_per = int(np.round(256 * 0.25))
    gt_img = gt_classes[:, 0] - image_offset
    all_boxes = np.concatenate([rois[:, 1:], gt_boxes], 0).astype(np.float32)
    ims = np.concatenate([rois[:, 0].astype(np.int64), gt_img], 0)
    idx = np.argsort(ims, kind="stable") if order is None else np.asarray(order)
    im_sorted, all_boxes = ims[idx], all_boxes[idx]
    out_r, out_l, out_t = [], [], []
    for im in range(int(im_sorted[-1]) + 1):
        g = np.where(gt_img == im)[0]
        if g.size == 0:
            continue
        gs, ge = g[0], g[-1] + 1
        t = np.where(im_sorted == im)[0]
        ts, te = t[0], t[-1] + 1
        ious = ops.bbox_overlaps_f32(all_boxes[ts:te], gt_boxes[gs:ge])
        mo = ious.max(1)
        ga = ious.argmax(1) + gs
        fg = np.where(mo >= fg_thresh)[0]
        nfg = min(fg_per, fg.shape[0])
        if fg.size > 0:
            fg = rng.choice(fg, size=nfg, replace=False)
        bgi = np.where((mo < 0.5) & (mo >= 0.0))[0]
        nbg = min(256 - nfg, bgi.size)
        if bgi.size > 0:
            bgi = rng.choice(bgi, size=nbg, replace=False)
        keep = np.append(fg, bgi).astype(np.int64)
        if keep.size == 0:
            continue
        lab = gt_classes[:, 1][ga[keep]].copy()
        if nfg < lab.shape[0]:
            lab[nfg:] = 0
        out_r.append(np.column_stack((im_sorted[ts:te][keep].astype(np.float32), all_boxes[ts:te][keep])))
        out_l.append(lab)
        out_t.append(gt_boxes[ga[keep]])
    return np.concatenate(out_r, 0), np.concatenate(out_l, 0), np.concatenate(out_t, 0)
