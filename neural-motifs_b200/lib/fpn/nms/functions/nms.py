"""Greedy NMS operator over the on-device sm_100a kernels (csrc/nms.cu).

Same surface as the reference's lib/fpn/nms/functions/nms.py:7-45:
`apply_nms(scores, boxes, pre_nms_topn, post_nms_topn, boxes_per_im=None, nms_thresh)`
-> LongTensor of kept indices (score order within each image), or (indices, per-image counts).
Non-differentiable; plain CUDA tensors.

Unlike the reference (per image: sort, cudaMalloc, mask kernel, 4.5 MB D2H, CPU loop,
cudaFree), all images are reduced by ONE segmented launch pair and only the per-image keep
counts (a few ints) are read back, because the return type is a Python list.
"""
import numpy as np
import torch

import motifs_cabi as _c


def nms_segments(boxes_sorted, seg_sizes, thresh, max_keep=None):
    """Segmented greedy NMS. boxes_sorted [total,4] fp32 CUDA, each segment sorted by
    descending score; seg_sizes: python list of ints. Returns (keep [total] int32 with the
    kept LOCAL indices of segment s at offset seg_off[s], num_keep [S] int32, seg_off list)."""
    _c.require_cuda(boxes_sorted)
    dev = boxes_sorted.device
    S = len(seg_sizes)
    sizes = np.asarray(seg_sizes, dtype=np.int64)
    seg_off = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(sizes, out=seg_off[1:])
    words = sizes * ((sizes + 63) // 64)
    mask_off = np.zeros(S, dtype=np.int64)
    if S > 1:
        np.cumsum(words[:-1], out=mask_off[1:])
    total = int(seg_off[-1])
    max_seg = int(sizes.max()) if S else 0
    keep = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    num_keep = torch.zeros(max(S, 1), dtype=torch.int32, device=dev)
    if S == 0 or total == 0:
        return keep[:total], num_keep[:S], seg_off
    seg_off_d = torch.from_numpy(seg_off.astype(np.int32)).to(dev, non_blocking=True)
    mask_off_d = torch.from_numpy(mask_off).to(dev, non_blocking=True)
    mask = torch.empty(max(int(words.sum()), 1), dtype=torch.int64, device=dev)
    if max_keep is None:
        max_keep = max_seg
    lib = _c.load()
    with torch.cuda.device(dev):
        rc = lib.mb200_nms_segmented(_c.ptr(boxes_sorted), _c.ptr(seg_off_d), _c.ptr(mask_off_d), S, max_seg,
                                     float(thresh), int(max_keep), _c.ptr(mask), _c.ptr(keep), _c.ptr(num_keep),
                                     _c.cur_stream())
    _c.check(rc, "mb200_nms_segmented")
    return keep, num_keep, seg_off


def apply_nms(scores, boxes, pre_nms_topn=12000, post_nms_topn=2000, boxes_per_im=None, nms_thresh=0.7):
    _c.require_cuda(scores, boxes)
    just_inds = boxes_per_im is None
    if boxes_per_im is None:
        boxes_per_im = [boxes.size(0)]
    boxes_per_im = [int(b) for b in boxes_per_im]
    boxes = boxes.contiguous().float()
    # per image: descending sort, truncate to pre_nms_topn (nms.py:37-40)
    idx_list, sizes = [], []
    s = 0
    for bpi in boxes_per_im:
        e = s + bpi
        _, idx = torch.sort(scores[s:e], dim=0, descending=True)
        if idx.size(0) > pre_nms_topn:
            idx = idx[:pre_nms_topn]
        idx_list.append(idx + s)
        sizes.append(int(idx.size(0)))
        s = e
    order = torch.cat(idx_list, 0) if idx_list else boxes.new_zeros(0, dtype=torch.long)
    boxes_sorted = boxes[order].contiguous()
    keep, num_keep, seg_off = nms_segments(boxes_sorted, sizes, nms_thresh)
    counts = num_keep.cpu().tolist() if len(sizes) else []   # the only host read (list return type)
    out, im_per = [], []
    for i, n in enumerate(counts):
        n = min(int(n), post_nms_topn)
        o = int(seg_off[i])
        out.append(order[o + keep[o:o + n].long()])
        im_per.append(n)
    inds = torch.cat(out, 0) if out else order[:0]
    if just_inds:
        return inds
    return inds, im_per
