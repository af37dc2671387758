"""CPU: the product's host-side logic (index math, packing, sorting keys, sampling) against the golden
fixtures made by the reference's own code and against the oracle — no kernels involved."""
import numpy as np
import pytest
import torch

from oracle import host as OH
from oracle import model as OM


def test_generate_anchors_matches_reference(golden):
    from lib.fpn.generate_anchors import generate_anchors
    from config import ANCHOR_SIZE, ANCHOR_SCALES, ANCHOR_RATIOS
    a = generate_anchors(base_size=ANCHOR_SIZE, feat_stride=16, anchor_scales=ANCHOR_SCALES, anchor_ratios=ANCHOR_RATIOS)
    assert np.array_equal(a, golden["anchors"])


def test_packing_helpers_match_reference(golden):
    from lib.pytorch_misc import transpose_packed_sequence_inds, enumerate_by_image, to_onehot, gather_nd, diagonal_inds
    inds, lens = transpose_packed_sequence_inds([int(x) for x in golden["tp_lengths"]])
    assert np.array_equal(inds, golden["tp_inds"]) and np.array_equal(lens, golden["tp_lens"])
    assert np.array_equal(np.array(list(enumerate_by_image(torch.from_numpy(golden["ebi_in"])))), golden["ebi_out"])
    oh = to_onehot(torch.tensor([2, 0]), 4)
    assert oh.tolist() == [[-1000, -1000, 1000, -1000], [1000, -1000, -1000, -1000]]
    x = torch.arange(2 * 3 * 4 * 5).view(2, 3, 4, 5)
    idx = torch.tensor([[1, 2, 3], [0, 0, 1]])
    assert torch.equal(gather_nd(x, idx), torch.stack([x[1, 2, 3], x[0, 0, 1]]))
    assert diagonal_inds(torch.zeros(3, 3)).tolist() == [0, 4, 8]


def test_box_utils_torch_paths_match_reference(golden):
    from lib.fpn import box_utils
    b = torch.from_numpy(golden["bp_boxes"])
    assert np.array_equal(box_utils.center_size(b).numpy(), golden["center_size"])
    assert np.array_equal(box_utils.point_form(torch.from_numpy(golden["center_size"])).numpy(), golden["point_form"])
    assert np.array_equal(box_utils.nms_overlaps(torch.from_numpy(golden["nmsov_boxes"])).numpy(), golden["nmsov_out"])
    out = box_utils.bbox_preds(b, torch.from_numpy(golden["bp_deltas"]))       # CPU tensors: torch path
    np.testing.assert_allclose(out.numpy(), golden["bp_out"], rtol=1e-6, atol=1e-5)


def test_sort_by_score_matches_oracle():
    from lib.rel_model import _sort_by_score
    rng = np.random.RandomState(0)
    im = np.repeat(np.arange(5), [7, 3, 9, 1, 4])
    scores = torch.from_numpy(rng.rand(im.shape[0]).astype(np.float32))
    p, ip, ls = _sort_by_score(torch.from_numpy(im), scores)
    po, ipo, lso = OM.sort_by_score(torch.from_numpy(im), scores)
    assert torch.equal(p, po) and torch.equal(ip, ipo) and list(ls) == list(lso)
    # images come out longest first, each image's objects by descending score
    first = im[p.numpy()][:ls[0]]
    assert len(set(first.tolist())) == ls[0]
    # the caller-held host copy of the image indices (MOTIFS_EARLY_HOST_INDS) gives the same permutation
    p2, ip2, ls2 = _sort_by_score(torch.from_numpy(im), scores, host=im)
    assert torch.equal(p2, p) and torch.equal(ip2, ip) and list(ls2) == list(ls)


def test_proposal_assignments_gtbox_matches_oracle():
    from lib.fpn.proposal_assignments.proposal_assignments_gtbox import proposal_assignments_gtbox
    from dataloaders.synthetic import make_numpy_batch
    nb = make_numpy_batch(3, seed=5, boxes_per_img=20, rels_per_img=15, image_offset=6)
    gt_boxes = torch.from_numpy(nb["gt_boxes"]); gt_classes = torch.from_numpy(nb["gt_classes"])
    gt_rels = torch.from_numpy(nb["gt_rels"])
    rois = torch.cat(((gt_classes[:, 0] - 6).float()[:, None], gt_boxes), 1)
    _, labels, rel = proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, 6, rng=np.random.RandomState(3))
    exp = OM.proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, 6, np.random.RandomState(3))
    assert torch.equal(rel, exp) and rel.shape == (768, 4)
    _, _, rel2 = proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, 6, rng=np.random.RandomState(3), num_im=3)
    assert torch.equal(rel2, exp)
    assert torch.equal(labels, gt_classes[:, 1])
    assert int((rel[:, 3] > 0).sum()) == 45          # every GT relation kept (45 <= 0.25 * 256 * 3)


def test_synthetic_blob_contract():
    from dataloaders.synthetic import make_numpy_batch, to_tuple
    nb = make_numpy_batch(2, seed=0)
    t = to_tuple(nb, "cpu")
    assert len(t) == 8 and t[0].shape == (2, 3, 592, 592) and t[1].shape == (2, 3)
    assert t[3].shape == (40, 4) and t[4].shape == (40, 2) and t[5].shape == (30, 4)
    assert int(t[5][:, 1:3].max()) < 20 and (t[5][:, 1] != t[5][:, 2]).all()


def test_prefetch_loader_order_bounds_and_errors():
    """dataloaders/prefetch.PrefetchLoader, host logic on the CPU device: batches arrive in order
    with the forward tuple of Blob.__getitem__, the producer never runs more than `depth` + 1 batches ahead, an exception in
    the batch source surfaces in the consumer, and abandoning the iterator stops the thread."""
    import threading
    import time
    from dataloaders.prefetch import PrefetchLoader
    from dataloaders.synthetic import make_numpy_batch
    made = []

    def source(n, fail_at=None):
        for i in range(n):
            if fail_at == i:
                raise ValueError("bad record %d" % i)
            made.append(i)
            yield make_numpy_batch(1, seed=i, boxes_per_img=3, rels_per_img=2)

    seen = []
    for i, blob in enumerate(PrefetchLoader(source(6), "cpu", depth=2)):
        time.sleep(0.02)
        assert len(made) <= i + 1 + 2 + 1            # consumed + queue depth + the one being built
        blob.scatter()
        tup = blob[0]
        want = make_numpy_batch(1, seed=i, boxes_per_img=3, rels_per_img=2)
        assert np.array_equal(tup[0].numpy(), want["imgs"]) and np.array_equal(tup[4].numpy(), want["gt_classes"])
        assert tup[6] is None and tup[7] is None and tup[2] == 0
        seen.append(i)
    assert seen == list(range(6))
    with pytest.raises(ValueError, match="bad record 2"):
        for blob in PrefetchLoader(source(5, fail_at=2), "cpu", depth=1):
            pass
    before = threading.active_count()
    it = iter(PrefetchLoader(source(100), "cpu", depth=2))
    next(it); it.close()
    time.sleep(0.3)
    assert threading.active_count() <= before
