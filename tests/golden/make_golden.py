"""Generates tests/golden/*.npz by RUNNING THE REFERENCE's own code in this container:
  * its Cython modules compiled as is into oracle/_ref (oracle/Makefile),
  * its pure-Python helpers imported from /root/reference (box_utils, generate_anchors,
    anchor_targets, pytorch_misc) with `h5py` stubbed (not installed here).
The reference has no golden vectors of its own for this path; these fixtures are
the pin.  /root/reference does not exist on the GPU box, so only the committed .npz travel.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MOTIFS_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402


def import_reference():
    """Make `lib.*` / `config` resolve to the reference tree, with its two Cython modules served
    from oracle/_ref and h5py stubbed."""
    sys.path.insert(0, REF)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import lib  # reference's package
    import lib.fpn  # noqa
    bbox = ref_loader.ref_bbox()
    draw = ref_loader.ref_draw_rectangles()
    assert bbox is not None and draw is not None, "run `make -C oracle` first"
    pkg = types.ModuleType("lib.fpn.box_intersections_cpu")
    pkg.__path__ = []
    sys.modules["lib.fpn.box_intersections_cpu"] = pkg
    sys.modules["lib.fpn.box_intersections_cpu.bbox"] = bbox
    pkg2 = types.ModuleType("lib.draw_rectangles")
    pkg2.__path__ = []
    sys.modules["lib.draw_rectangles"] = pkg2
    sys.modules["lib.draw_rectangles.draw_rectangles"] = draw
    return bbox, draw


def rand_boxes(rng, n, lo=1.0, hi=190.0, size=592):
    x1 = rng.uniform(0, 400, n)
    y1 = rng.uniform(0, 400, n)
    w = rng.uniform(lo, hi, n)
    h = rng.uniform(lo, hi, n)
    return np.stack([x1, y1, np.minimum(x1 + w, size - 1), np.minimum(y1 + h, size - 1)], 1).astype(np.float32)


def main():
    bbox, draw = import_reference()
    import torch
    from lib.fpn import box_utils as ref_box_utils
    from lib.fpn.generate_anchors import generate_anchors
    from lib.fpn.anchor_targets import anchor_target_layer
    from lib.pytorch_misc import transpose_packed_sequence_inds, enumerate_by_image
    import config as ref_config

    rng = np.random.RandomState(1234)
    g = {}
    # ---- Cython: float64 IoU / intersections (bbox.pyx) incl. degenerate + identical + disjoint boxes
    a = rand_boxes(rng, 257)
    b = rand_boxes(rng, 41)
    a[0] = b[0]                      # identical
    a[1] = [0, 0, 0, 0]              # 1-pixel box
    b[1] = [580, 580, 591, 591]      # far corner
    g["iou_a"], g["iou_b"] = a, b
    g["iou_f64"] = bbox.bbox_overlaps(a.astype(np.float64), b.astype(np.float64))
    g["inter_f64"] = bbox.bbox_intersections(a.astype(np.float64), b.astype(np.float64))
    # ---- Cython: rasteriser (draw_rectangles.pyx), P = 27 and 13
    pairs = np.concatenate([rand_boxes(rng, 300), rand_boxes(rng, 300)], 1)
    pairs[0, 4:] = pairs[0, :4]      # identical pair
    g["draw_pairs"] = pairs
    g["draw_27"] = draw.draw_union_boxes(pairs, 27)
    g["draw_13"] = draw.draw_union_boxes(pairs[:50], 13)
    # ---- box_utils (torch CPU path of the reference)
    boxes = rand_boxes(rng, 500)
    deltas = (rng.randn(500, 4) * np.array([0.1, 0.1, 0.2, 0.2])).astype(np.float32)
    g["bp_boxes"], g["bp_deltas"] = boxes, deltas
    g["bp_out"] = ref_box_utils.bbox_preds(torch.from_numpy(boxes), torch.from_numpy(deltas)).numpy()
    g["center_size"] = ref_box_utils.center_size(torch.from_numpy(boxes)).numpy()
    g["point_form"] = ref_box_utils.point_form(torch.from_numpy(g["center_size"])).numpy()
    g["iou_f32"] = ref_box_utils.bbox_overlaps(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    cls_boxes = np.stack([rand_boxes(rng, 24) for _ in range(5)], 1)  # [24,5,4]
    g["nmsov_boxes"] = cls_boxes
    g["nmsov_out"] = ref_box_utils.nms_overlaps(torch.from_numpy(cls_boxes)).numpy()
    # ---- anchors
    ans = generate_anchors(base_size=ref_config.ANCHOR_SIZE, feat_stride=16,
                           anchor_scales=ref_config.ANCHOR_SCALES, anchor_ratios=ref_config.ANCHOR_RATIOS)
    g["anchors"] = ans
    # ---- anchor targets: the deterministic part (labels before subsampling) is recovered by
    # seeding numpy's global RNG exactly as the test will.
    gt = rand_boxes(rng, 12, lo=32, hi=300).astype(np.float32)
    np.random.seed(7)
    anchors, anchor_inds, bbox_targets, labels = anchor_target_layer(gt, (592, 592))
    g["at_gt"], g["at_anchors"], g["at_inds"], g["at_targets"], g["at_labels"] = gt, anchors, anchor_inds, bbox_targets, labels
    # ---- packing helpers
    lengths = [9, 7, 7, 4, 1]
    inds, lens = transpose_packed_sequence_inds(lengths)
    g["tp_lengths"], g["tp_inds"], g["tp_lens"] = np.array(lengths), np.asarray(inds), np.array(lens)
    im_inds = torch.LongTensor([0, 0, 0, 1, 1, 3, 3, 3, 3])
    g["ebi_in"] = im_inds.numpy()
    g["ebi_out"] = np.array(list(enumerate_by_image(im_inds)))
    np.savez_compressed(os.path.join(OUT, "reference_host_ops.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_host_ops.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
