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
