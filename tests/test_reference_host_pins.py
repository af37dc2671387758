"""Pins of the oracle restatements — and of the product's host functions that run on CPU tensors — against fixtures
produced by RUNNING THE REFERENCE's own code in the build container (tests/golden/make_golden_host2.py ->
reference_host_ops2.npz): proposal_assignments_gtbox / _det with the numpy RNG consumed in the reference's order,
surgery.filter_dets, rel_model._sort_by_score (the reference holds no golden vectors for this path,
so outputs of the reference itself are the pin)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_host_ops2.npz"))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_proposal_assignments_gtbox_oracle_and_product_match_reference(g2, tag):
    from oracle import model as OM
    from lib.fpn.proposal_assignments.proposal_assignments_gtbox import proposal_assignments_gtbox
    off, seed = [int(v) for v in g2["gtbox_%s_meta" % tag]]
    rois, gt_boxes, gt_classes, gt_rels = (torch.from_numpy(g2["gtbox_%s_%s" % (tag, k)])
                                           for k in ("rois", "gt_boxes", "gt_classes", "gt_rels"))
    want = torch.from_numpy(g2["gtbox_%s_rel_labels" % tag])
    got_o = OM.proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, off, np.random.RandomState(seed))
    assert torch.equal(got_o, want)
    _, labels, got_p = proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, off, rng=np.random.RandomState(seed))
    assert torch.equal(got_p, want)
    assert np.array_equal(labels.numpy(), g2["gtbox_%s_labels" % tag])


def test_proposal_assignments_det_oracle_matches_reference(g2):
    """Exact match once the oracle is given the candidate order the reference's (unstable) torch.sort produced; with
    its own stable order the same rule set applies (same counts of foreground / background rows per image)."""
    from oracle import host
    off, seed = [int(v) for v in g2["det_meta"]]
    args = (g2["det_rois"], g2["det_gt_boxes"], g2["det_gt_classes"], off)
    r, l, t = host.proposal_assignments_det(*args, np.random.RandomState(seed), order=g2["det_sort_idx"])
    assert np.array_equal(r, g2["det_out_rois"]) and np.array_equal(l, g2["det_out_labels"])
    assert np.array_equal(t, g2["det_out_targets"])
    r2, l2, t2 = host.proposal_assignments_det(*args, np.random.RandomState(seed))
    assert r2.shape == r.shape
    for im in range(2):
        assert int(((r2[:, 0] == im) & (l2 > 0)).sum()) == int(((r[:, 0] == im) & (l > 0)).sum())


def test_filter_dets_order_matches_reference(g2):
    """surgery.py:21-59: relations sorted by (best non-background predicate) x obj score x obj score, descending."""
    obj_scores, pred_scores, rel_inds = g2["fd_in_obj_scores"], g2["fd_in_pred_scores"], g2["fd_in_rel_inds"]
    key = pred_scores[:, 1:].max(1) * obj_scores[rel_inds[:, 0]] * obj_scores[rel_inds[:, 1]]
    order = np.argsort(-key, kind="stable")
    assert np.array_equal(rel_inds[order], g2["fd_out_rels"])
    assert np.array_equal(pred_scores[order], g2["fd_out_pred_scores"])
    assert np.array_equal(g2["fd_out_boxes"], g2["fd_in_boxes"]) and np.array_equal(g2["fd_out_objs"], g2["fd_in_obj_classes"])
    # the oracle's inlined version (oracle/model.py, end of RelModel.forward) is the same expression in torch
    s = torch.from_numpy(obj_scores); ri = torch.from_numpy(rel_inds); rr = torch.from_numpy(pred_scores)
    _, idx = torch.sort((rr[:, 1:].max(1)[0] * s[ri[:, 0]] * s[ri[:, 1]]).view(-1), dim=0, descending=True, stable=True)
    assert np.array_equal(ri[idx].numpy(), g2["fd_out_rels"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_sort_by_score_oracle_and_product_match_reference(g2, tag):
    from oracle import model as OM
    from lib.rel_model import _sort_by_score
    im = torch.from_numpy(g2["sort_%s_im" % tag]); scores = torch.from_numpy(g2["sort_%s_scores" % tag])
    for fn in (OM.sort_by_score, _sort_by_score):
        perm, inv, ls = fn(im, scores)
        assert list(ls) == g2["sort_%s_ls" % tag].tolist()
        if tag == "d":      # two equal scores inside an image: the reference's torch.sort leaves their order unspecified
            want = g2["sort_%s_perm" % tag]
            assert np.array_equal(scores.numpy()[perm.numpy()], scores.numpy()[want])
            assert np.array_equal(im.numpy()[perm.numpy()], im.numpy()[want])
            assert np.array_equal(np.sort(perm.numpy()), np.arange(len(want)))
            continue
        assert np.array_equal(inv.numpy(), g2["sort_%s_inv" % tag]) and np.array_equal(perm.numpy(), g2["sort_%s_perm" % tag])
    perm, inv, ls = _sort_by_score(im, scores, host=g2["sort_%s_im" % tag])
    assert np.array_equal(scores.numpy()[perm.numpy()], scores.numpy()[g2["sort_%s_perm" % tag]])


def test_decoder_rnn_oracle_matches_reference_module():
    """oracle/model.py:DecoderRNN against the REFERENCE's lib/lstm/decoder_rnn.py run on the CPU
    (tests/golden/make_golden_decoder.py): teacher-forced training forward with background labels, and greedy eval
    with the overlap-aware commitment loop (the fixture's commitments differ from plain greedy ones)."""
    from oracle import model as OM
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_decoder.npz"))
    H, D = [int(v) for v in g["dims"]]
    classes = ['__background__'] + ['c%d' % i for i in range(150)]
    dec = OM.DecoderRNN(classes, D, H)
    dec.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")})
    dec.train()
    dists, commits = dec(torch.from_numpy(g["tr_x"]), g["tr_bl"].tolist(), labels=torch.from_numpy(g["tr_labels"]))
    assert np.array_equal(commits.numpy(), g["tr_commits"])
    assert np.allclose(dists.detach().numpy(), g["tr_dists"], rtol=1e-5, atol=1e-6)
    dec.eval()
    T = g["ev_x"].shape[0]
    with torch.no_grad():
        dists, commits = dec(torch.from_numpy(g["ev_x"]), [1] * T, boxes_for_nms=torch.from_numpy(g["ev_boxes"]))
        _, greedy = dec(torch.from_numpy(g["ev_x"]), [1] * T)
    assert np.allclose(dists.numpy(), g["ev_dists"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(commits.numpy(), g["ev_commits"])
    assert np.array_equal(greedy.numpy(), g["ev_commits_greedy"])
    assert not np.array_equal(g["ev_commits"], g["ev_commits_greedy"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_rel_assignments_oracle_and_product_match_reference(tag):
    """SGDet-training relation labels (rel_assignments.py:15-145) with the numpy RNG consumed in the reference's
    order: foreground sampling proportional to IoU products, background sampling, per-image lexsort."""
    from oracle import host
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_rel_assignments.npz"))
    off, seed, nsg, fno = [int(v) for v in g["ra_%s_meta" % tag]]
    a = {k: g["ra_%s_%s" % (tag, k)] for k in ("ims", "boxes", "labels", "gt_boxes", "gt_classes", "gt_rels", "out")}
    got = host.rel_assignments(a["ims"], a["boxes"], a["labels"], a["gt_boxes"], a["gt_classes"], a["gt_rels"], off,
                               np.random.RandomState(seed), num_sample_per_gt=nsg, filter_non_overlap=bool(fno))
    assert np.array_equal(got, a["out"])
    assert int((a["out"][:, 3] > 0).sum()) > 0            # the fixture holds foreground relations


def test_forward_tuple_layout_matches_reference_blob():
    """dataloaders/synthetic.py hands `RelModel.forward` the same positional tuple the reference's Blob produces from
    the same images / boxes / relations (blob.py:62-229 run on the CPU, tests/golden/make_golden_blob.py): gt_classes
    rows = (image index within the batch, class), gt_rels rows = (image, subject, object, predicate) with box indices
    local to the image, im_sizes rows = (h, w, scale), image_offset 0 on a single GPU; and the training tuple's
    train_anchor_inds = (image, h, w, anchor) rows of anchor_target_layer, per image in order."""
    from dataloaders.synthetic import to_tuple, SyntheticBlob
    from oracle.host import anchor_target_layer        # the product's runs its IoU on the GPU (parity-tested there)
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_blob.npz"))
    nb = dict(imgs=g["nb_imgs"], im_sizes=g["nb_im_sizes"], image_offset=0, gt_boxes=g["nb_gt_boxes"],
              gt_classes=g["nb_gt_classes"], gt_rels=g["nb_gt_rels"])
    for tup in (to_tuple(nb, "cpu"), (lambda b: (b.scatter(), b[0])[1])(SyntheticBlob(nb, "cpu"))):
        assert len(tup) == 8 and tup[6] is None
        for tag in ("train", "eval"):
            assert np.array_equal(tup[0].numpy(), g[tag + "_imgs"])
            assert np.array_equal(np.asarray(tup[1], dtype=np.float64), np.asarray(g[tag + "_im_sizes"], dtype=np.float64))
            assert int(tup[2]) == int(g[tag + "_image_offset"]) == 0
            assert np.array_equal(tup[3].numpy(), g[tag + "_gt_boxes"]) and tup[3].dtype == torch.float32
            assert np.array_equal(tup[4].numpy(), g[tag + "_gt_classes"]) and tup[4].dtype == torch.int64
            assert np.array_equal(tup[5].numpy(), g[tag + "_gt_rels"]) and tup[5].dtype == torch.int64
    # train_anchor_inds: the reference calls anchor_target_layer per appended image with the global numpy RNG
    np.random.seed(5)
    rows = []
    for i in range(3):
        gb = nb["gt_boxes"][nb["gt_classes"][:, 0] == i]
        h, w = nb["im_sizes"][i][:2]
        _, inds, _, _ = anchor_target_layer(gb, (h, w), rng=np.random)
        rows.append(np.column_stack((np.full(inds.shape[0], i, dtype=np.int64), inds)))
    assert np.array_equal(np.concatenate(rows, 0), g["train_anchor_inds"])


@pytest.mark.parametrize("tag,max_norm", [("clip", 5.0), ("noclip", 1e6)])
def test_clip_grad_norm_matches_reference(g2, tag, max_norm):
    """lib/pytorch_misc.clip_grad_norm (one device reduction) against the reference's per-parameter loop (:416-459)."""
    from lib.pytorch_misc import clip_grad_norm
    ps = []
    for i in range(4):
        gr = g2["cgn_grad%d" % i]
        p = torch.nn.Parameter(torch.zeros(gr.shape)); p.grad = torch.from_numpy(gr.copy()); ps.append(("p%d" % i, p))
    ps.append(("nograd", torch.nn.Parameter(torch.zeros(4))))
    total = clip_grad_norm(ps, max_norm=max_norm, clip=True)
    assert abs(float(total) - float(g2["cgn_%s_total" % tag])) < 1e-5 * float(g2["cgn_%s_total" % tag])
    for i in range(4):
        assert np.allclose(ps[i][1].grad.numpy(), g2["cgn_%s_after%d" % (tag, i)], rtol=1e-6, atol=1e-7)
