"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's native operators.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; the product (`neural-motifs_b200/`) never does.

Parity pin: the reference holds no golden vectors for this path. The oracle is
pinned instead against (a) the reference's own Cython modules (`bbox.pyx`,
`draw_rectangles.pyx`) and pure-Python helpers (`box_utils.py`, `generate_anchors.py`) run
HERE from /root/reference — fixtures in tests/golden/ made by tests/golden/make_golden.py — and
(b) on the GPU box, against the reference's CUDA kernels compiled unmodified for sm_100a
(oracle/_ref/libref_kernels.so, built by oracle/Makefile).

Floating-point contract: where the reference's CUDA source leaves the compiler free to fuse a
multiply-add, the restatement uses the fusion nvcc 12.9 actually emits for that source on
sm_100a (read from its SASS; DESIGN.md "fp contract"), emulated exactly through float64.
"""
import numpy as np

f32 = np.float32


def _fma32(a, b, c):
    """fp32 fused multiply-add: exact product and sum in float64, one rounding to fp32
    (the float64 sum can double-round only in ~2^-29 of cases)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


# --------------------------------------------------------------------------- RoIAlign
def normalize_rois(rois, feat_h, feat_w, spatial_scale):
    """lib/fpn/roi_align/functions/roi_align.py:20-31 (fp32, in place on a clone)."""
    r = np.array(rois, dtype=f32, copy=True)
    height = f32((feat_h - 1) / spatial_scale)
    width = f32((feat_w - 1) / spatial_scale)
    r[:, 1] /= width
    r[:, 2] /= height
    r[:, 3] /= width
    r[:, 4] /= height
    return r


def _axis_samples(a1, a2, size, crop):
    """roi_align_kernel.cu:37-55 for one axis, vectorised over rois. Returns in [N,crop], ok."""
    a1 = a1.astype(f32)
    a2 = a2.astype(f32)
    n = a1.shape[0]
    if crop > 1:
        scale = ((a2 - a1) * f32(size - 1)).astype(f32) / f32(crop - 1)
        scale = scale.astype(f32)
        i = np.arange(crop, dtype=f32)[None, :]
        prod = (i * scale[:, None]).astype(f32)
        pos = _fma32(a1[:, None], f32(size - 1), prod)
    else:
        pos = (0.5 * (a1 + a2).astype(np.float64) * (size - 1)).astype(f32)[:, None]
    ok = ~((pos < 0) | (pos > f32(size - 1)))
    return pos, ok


def roi_align_forward(features, boxes_norm, crop_h, crop_w, extrapolation_value=0.0):
    """roi_align_kernel.cu:15-80. features [B,C,H,W] fp32, boxes_norm [N,5] normalised.
    Returns [N,C,crop_h,crop_w]; rois with an out-of-range batch index give zeros (the
    caller's zero fill, functions/roi_align.py:36)."""
    features = np.asarray(features, f32)
    boxes = np.asarray(boxes_norm, f32)
    B, C, H, W = features.shape
    N = boxes.shape[0]
    out = np.zeros((N, C, crop_h, crop_w), f32)
    if N == 0:
        return out
    b_in = boxes[:, 0].astype(np.int32)  # int() truncation
    in_y, ok_y = _axis_samples(boxes[:, 2], boxes[:, 4], H, crop_h)
    in_x, ok_x = _axis_samples(boxes[:, 1], boxes[:, 3], W, crop_w)
    for n in range(N):
        b = int(b_in[n])
        if b < 0 or b >= B:
            continue
        iy = np.where(ok_y[n], in_y[n], 0).astype(f32)
        ix = np.where(ok_x[n], in_x[n], 0).astype(f32)
        ty = np.floor(iy).astype(np.int64)
        by = np.ceil(iy).astype(np.int64)
        ly = (iy - ty.astype(f32)).astype(f32)
        lx_i = np.floor(ix).astype(np.int64)
        rx_i = np.ceil(ix).astype(np.int64)
        lx = (ix - lx_i.astype(f32)).astype(f32)
        img = features[b]  # [C,H,W]
        tl = img[:, ty][:, :, lx_i]
        tr = img[:, ty][:, :, rx_i]
        bl = img[:, by][:, :, lx_i]
        br = img[:, by][:, :, rx_i]
        wx = lx[None, None, :]
        wy = ly[None, :, None]
        top = _fma32(wx, (tr - tl).astype(f32), tl)
        bottom = _fma32(wx, (br - bl).astype(f32), bl)
        val = _fma32(wy, (bottom - top).astype(f32), top)
        ok = ok_y[n][:, None] & ok_x[n][None, :]
        out[n] = np.where(ok[None], val, f32(extrapolation_value))
    return out


def roi_align_backward(grads, boxes_norm, batch, channels, H, W):
    """roi_align_kernel.cu:103-170 (4 scatter-adds per element; float64 accumulation here so the
    comparison with the atomics' arbitrary order is tolerance based)."""
    grads = np.asarray(grads, f32)
    boxes = np.asarray(boxes_norm, f32)
    N, C, PH, PW = grads.shape
    out = np.zeros((batch, channels, H, W), np.float64)
    if N == 0:
        return out.astype(f32)
    b_in = boxes[:, 0].astype(np.int32)
    in_y, ok_y = _axis_samples(boxes[:, 2], boxes[:, 4], H, PH)
    in_x, ok_x = _axis_samples(boxes[:, 1], boxes[:, 3], W, PW)
    for n in range(N):
        b = int(b_in[n])
        if b < 0 or b >= batch:
            continue
        for y in range(PH):
            if not ok_y[n, y]:
                continue
            iy = in_y[n, y]
            ty, by = int(np.floor(iy)), int(np.ceil(iy))
            wy = f32(iy - f32(ty))
            for x in range(PW):
                if not ok_x[n, x]:
                    continue
                ix = in_x[n, x]
                lx, rx = int(np.floor(ix)), int(np.ceil(ix))
                wx = f32(ix - f32(lx))
                g = grads[n, :, y, x]
                dtop = (f32(1) - wy) * g
                dbot = wy * g
                out[b, :, ty, lx] += (f32(1) - wx) * dtop
                out[b, :, ty, rx] += wx * dtop
                out[b, :, by, lx] += (f32(1) - wx) * dbot
                out[b, :, by, rx] += wx * dbot
    return out.astype(f32)


# --------------------------------------------------------------------------- NMS
def dev_iou_matrix(boxes_a, boxes_b):
    """devIoU, nms_kernel.cu:23-31, fp32, a = row box (cur_box), b = column box, with the
    contraction nvcc emits: Sa = mul, Sa+Sb = fma(wb, hb, Sa), inter = mul, IEEE divide."""
    a = np.asarray(boxes_a, f32)[:, None, :]
    b = np.asarray(boxes_b, f32)[None, :, :]
    one = f32(1)
    left = np.maximum(a[..., 0], b[..., 0])
    right = np.minimum(a[..., 2], b[..., 2])
    top = np.maximum(a[..., 1], b[..., 1])
    bottom = np.minimum(a[..., 3], b[..., 3])
    width = np.maximum(((right - left).astype(f32) + one).astype(f32), f32(0))
    height = np.maximum(((bottom - top).astype(f32) + one).astype(f32), f32(0))
    inter = (width * height).astype(f32)
    sa = (((a[..., 2] - a[..., 0]).astype(f32) + one).astype(f32) *
          ((a[..., 3] - a[..., 1]).astype(f32) + one).astype(f32)).astype(f32)
    wb = ((b[..., 2] - b[..., 0]).astype(f32) + one).astype(f32)
    hb = ((b[..., 3] - b[..., 1]).astype(f32) + one).astype(f32)
    sasb = _fma32(wb, hb, sa)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (sasb - inter).astype(f32)).astype(f32)


def nms_keep(boxes_sorted, thresh):
    """ApplyNMSGPU, nms_kernel.cu:88-131: greedy scan in index order over boxes already sorted
    by score; box j is suppressed by a kept box i<j when devIoU(i, j) > thresh (strict)."""
    boxes = np.asarray(boxes_sorted, f32)
    n = boxes.shape[0]
    removed = np.zeros(n, bool)
    keep = []
    thresh = f32(thresh)
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if i + 1 < n:
            iou = dev_iou_matrix(boxes[i:i + 1], boxes[i + 1:])[0]
            removed[i + 1:] |= iou > thresh
    return np.asarray(keep, np.int32)


def apply_nms(scores, boxes, pre_nms_topn=12000, post_nms_topn=2000, boxes_per_im=None, nms_thresh=0.7):
    """lib/fpn/nms/functions/nms.py:7-45 on numpy arrays. Returns int64 indices (and per-image
    counts when boxes_per_im is given). Sorting is a stable descending sort: the reference's
    torch.sort makes no tie guarantee, so parity inputs must not contain tied scores."""
    scores = np.asarray(scores, f32)
    boxes = np.asarray(boxes, f32)
    just_inds = boxes_per_im is None
    if boxes_per_im is None:
        boxes_per_im = [boxes.shape[0]]
    s = 0
    keep, im_per = [], []
    for bpi in boxes_per_im:
        e = s + int(bpi)
        idx = np.argsort(-scores[s:e], kind="stable")
        if idx.shape[0] > pre_nms_topn:
            idx = idx[:pre_nms_topn]
        k = nms_keep(boxes[s:e][idx], nms_thresh)
        k = k[:min(k.shape[0], post_nms_topn)]
        keep.append(idx[k].astype(np.int64) + s)
        im_per.append(int(k.shape[0]))
        s = e
    inds = np.concatenate(keep, 0) if keep else np.zeros(0, np.int64)
    if just_inds:
        return inds
    return inds, im_per


# --------------------------------------------------------------------------- box utilities
def center_size(boxes):
    """lib/fpn/box_utils.py:51-63."""
    boxes = np.asarray(boxes, f32)
    wh = (boxes[:, 2:] - boxes[:, :2] + f32(1.0)).astype(f32)
    return np.column_stack(((boxes[:, :2] + f32(0.5) * wh).astype(f32), wh)).astype(f32)


def point_form(boxes):
    """lib/fpn/box_utils.py:66-79."""
    boxes = np.asarray(boxes, f32)
    return np.column_stack(((boxes[:, :2] - f32(0.5) * boxes[:, 2:]).astype(f32),
                            (boxes[:, :2] + f32(0.5) * (boxes[:, 2:] - f32(2.0))).astype(f32))).astype(f32)


def bbox_preds(boxes, deltas):
    """lib/fpn/box_utils.py:28-48."""
    boxes = np.asarray(boxes, f32)
    deltas = np.asarray(deltas, f32)
    if boxes.shape[0] == 0:
        return boxes
    pc = center_size(boxes)
    xys = (pc[:, :2] + (pc[:, 2:] * deltas[:, :2]).astype(f32)).astype(f32)
    whs = (np.exp(deltas[:, 2:]).astype(f32) * pc[:, 2:]).astype(f32)
    return point_form(np.concatenate((xys, whs), 1))


def bbox_overlaps_f32(box_a, box_b):
    """lib/fpn/box_utils.py:85-131 (torch branch), fp32, no fused ops."""
    a = np.asarray(box_a, f32)[:, None, :]
    b = np.asarray(box_b, f32)[None, :, :]
    one = f32(1.0)
    max_xy = np.minimum(a[..., 2:], b[..., 2:])
    min_xy = np.maximum(a[..., :2], b[..., :2])
    inter = np.maximum(((max_xy - min_xy).astype(f32) + one).astype(f32), f32(0))
    inter = (inter[..., 0] * inter[..., 1]).astype(f32)
    area_a = (((a[..., 2] - a[..., 0]).astype(f32) + one) * ((a[..., 3] - a[..., 1]).astype(f32) + one)).astype(f32)
    area_b = (((b[..., 2] - b[..., 0]).astype(f32) + one) * ((b[..., 3] - b[..., 1]).astype(f32) + one)).astype(f32)
    union = ((area_a + area_b).astype(f32) - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(f32)


def bbox_overlaps_f64(boxes, query_boxes, mode=0):
    """lib/fpn/box_intersections_cpu/bbox.pyx:21-62 (mode 0, IoU) and :71-107 (mode 1,
    intersection over query-box area); float64, zero unless iw > 0 and ih > 0."""
    b = np.ascontiguousarray(boxes, np.float64)[:, None, :]
    q = np.ascontiguousarray(query_boxes, np.float64)[None, :, :]
    box_area = (q[..., 2] - q[..., 0] + 1) * (q[..., 3] - q[..., 1] + 1)
    iw = np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0]) + 1
    ih = np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1]) + 1
    inter = iw * ih
    if mode == 0:
        ua = (b[..., 2] - b[..., 0] + 1) * (b[..., 3] - b[..., 1] + 1) + box_area - inter
        val = inter / ua
    else:
        val = inter / box_area + 0 * iw
    return np.where((iw > 0) & (ih > 0), val, 0.0)


def nms_overlaps(boxes):
    """lib/fpn/box_utils.py:134-154: per-class pairwise IoU, boxes [N,nc,4] -> [N,N,nc]."""
    boxes = np.asarray(boxes, f32)
    one = f32(1.0)
    max_xy = np.minimum(boxes[:, None, :, 2:], boxes[None, :, :, 2:])
    min_xy = np.maximum(boxes[:, None, :, :2], boxes[None, :, :, :2])
    inter = np.maximum(((max_xy - min_xy).astype(f32) + one).astype(f32), f32(0))
    inters = (inter[..., 0] * inter[..., 1]).astype(f32)
    areas = (((boxes[..., 2] - boxes[..., 0]).astype(f32) + one) *
             ((boxes[..., 3] - boxes[..., 1]).astype(f32) + one)).astype(f32)
    union = ((-inters + areas[None]).astype(f32) + areas[:, None]).astype(f32)
    return (inters / union).astype(f32)


# --------------------------------------------------------------------------- union boxes
def union_rois(rois, union_inds):
    """lib/get_union_boxes.py:82-87."""
    rois = np.asarray(rois, f32)
    a = rois[union_inds[:, 0]]
    b = rois[union_inds[:, 1]]
    return np.concatenate((a[:, :1], np.minimum(a[:, 1:3], b[:, 1:3]), np.maximum(a[:, 3:5], b[:, 3:5])), 1).astype(f32)


def draw_union_boxes(box_pairs, pooling_size):
    """lib/draw_rectangles/draw_rectangles.pyx:27-67, float32 arithmetic in source order."""
    p = np.asarray(box_pairs, f32)
    N = p.shape[0]
    P = int(pooling_size)
    out = np.zeros((N, 2, P, P), f32)
    if N == 0:
        return out
    x1u = np.minimum(p[:, 0], p[:, 4])
    y1u = np.minimum(p[:, 1], p[:, 5])
    x2u = np.maximum(p[:, 2], p[:, 6])
    y2u = np.maximum(p[:, 3], p[:, 7])
    w = (x2u - x1u).astype(f32)
    h = (y2u - y1u).astype(f32)
    Pf = f32(P)
    idx = np.arange(P, dtype=f32)

    def mm(x):
        return np.minimum(np.maximum(x, f32(0)), f32(1)).astype(f32)

    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(2):
            x1b = (((p[:, 0 + 4 * i] - x1u).astype(f32) * Pf).astype(f32) / w).astype(f32)
            y1b = (((p[:, 1 + 4 * i] - y1u).astype(f32) * Pf).astype(f32) / h).astype(f32)
            x2b = (((p[:, 2 + 4 * i] - x1u).astype(f32) * Pf).astype(f32) / w).astype(f32)
            y2b = (((p[:, 3 + 4 * i] - y1u).astype(f32) * Pf).astype(f32) / h).astype(f32)
            yc = (mm((idx[None] + f32(1)) - y1b[:, None]) * mm(y2b[:, None] - idx[None])).astype(f32)  # [N,P]
            xc = (mm((idx[None] + f32(1)) - x1b[:, None]) * mm(x2b[:, None] - idx[None])).astype(f32)
            out[:, i] = (xc[:, None, :] * yc[:, :, None]).astype(f32)
    return out
