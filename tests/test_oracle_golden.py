"""CPU: the oracle restatement against fixtures produced by the reference's OWN code
(tests/golden/make_golden.py: its Cython modules compiled as is, its Python helpers imported)."""
import numpy as np

from oracle import host, ops


def test_bbox_overlaps_f64_bit_exact(golden):
    assert np.array_equal(ops.bbox_overlaps_f64(golden["iou_a"], golden["iou_b"], 0), golden["iou_f64"])
    assert np.array_equal(ops.bbox_overlaps_f64(golden["iou_a"], golden["iou_b"], 1), golden["inter_f64"])


def test_draw_union_boxes_bit_exact(golden):
    assert np.array_equal(ops.draw_union_boxes(golden["draw_pairs"], 27), golden["draw_27"])
    assert np.array_equal(ops.draw_union_boxes(golden["draw_pairs"][:50], 13), golden["draw_13"])


def test_box_utils(golden):
    assert np.array_equal(ops.center_size(golden["bp_boxes"]), golden["center_size"])
    assert np.array_equal(ops.point_form(golden["center_size"]), golden["point_form"])
    # exp() differs by an ulp between numpy and torch: tolerance, not bits
    np.testing.assert_allclose(ops.bbox_preds(golden["bp_boxes"], golden["bp_deltas"]), golden["bp_out"],
                               rtol=1e-6, atol=1e-4)
    assert np.array_equal(ops.bbox_overlaps_f32(golden["iou_a"], golden["iou_b"]), golden["iou_f32"])
    assert np.array_equal(ops.nms_overlaps(golden["nmsov_boxes"]), golden["nmsov_out"])


def test_anchors_bit_exact(golden):
    assert np.array_equal(host.generate_anchors(), golden["anchors"])


def test_anchor_targets_with_same_rng(golden):
    np.random.seed(7)
    anchors, inds, targets, labels = host.anchor_target_layer(golden["at_gt"], (592, 592))
    assert np.array_equal(anchors, golden["at_anchors"])
    assert np.array_equal(inds, golden["at_inds"])
    assert np.array_equal(targets, golden["at_targets"])
    assert np.array_equal(labels, golden["at_labels"])


def test_packing_helpers(golden):
    inds, lens = host.transpose_packed_sequence_inds([int(x) for x in golden["tp_lengths"]])
    assert np.array_equal(inds, golden["tp_inds"]) and np.array_equal(lens, golden["tp_lens"])
    assert np.array_equal(np.array(host.enumerate_by_image(golden["ebi_in"])), golden["ebi_out"])


def test_nms_oracle_properties():
    """Greedy NMS invariants: kept boxes are mutually below threshold; every dropped box is
    above threshold with an earlier kept box; idempotent."""
    rng = np.random.RandomState(0)
    n = 400
    x1 = rng.uniform(0, 300, n); y1 = rng.uniform(0, 300, n)
    b = np.stack([x1, y1, x1 + rng.uniform(20, 200, n), y1 + rng.uniform(20, 200, n)], 1).astype(np.float32)
    keep = ops.nms_keep(b, 0.5)
    iou = ops.dev_iou_matrix(b, b)
    kk = iou[np.ix_(keep, keep)]
    assert (np.triu(kk, 1) <= np.float32(0.5)).all()
    dropped = np.setdiff1d(np.arange(n), keep)
    for d in dropped:
        earlier = keep[keep < d]
        assert (iou[earlier, d] > np.float32(0.5)).any()
    assert np.array_equal(ops.nms_keep(b[keep], 0.5), np.arange(len(keep)))
    assert ops.nms_keep(b[:0], 0.5).shape == (0,)


def test_roi_align_oracle_identity_and_extrapolation():
    """A box covering the whole map with crop == map size reproduces the map; samples outside
    the image take the extrapolation value; bad batch index gives zeros."""
    rng = np.random.RandomState(1)
    feat = rng.randn(2, 3, 5, 6).astype(np.float32)
    rois = np.array([[1, 0, 0, 1, 1], [0, -0.5, 0, 1, 1], [5, 0, 0, 1, 1]], np.float32)  # normalised
    out = ops.roi_align_forward(feat, rois, 5, 6, extrapolation_value=-7.0)
    np.testing.assert_allclose(out[0], feat[1], atol=1e-6)
    assert (out[1][:, :, 0] == -7.0).all()
    assert (out[2] == 0).all()
