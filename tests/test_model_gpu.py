"""Model-level parity on the GPU: the product RelModel (cuda:0, hand-written kernels) against the
CPU oracle restatement with the SAME state dict, inputs, dropout masks and sampling RNG.
Bar (north star): integer outputs (labels, relation triples, indices) identical; fp32 logits within
1e-3 of the oracle relative to the logits' max magnitude."""
import numpy as np
import pytest
import torch

from tests.model_utils import build_pair, make_masks, to_dev, relerr, l2err
from dataloaders.synthetic import make_numpy_batch, to_tuple

pytestmark = pytest.mark.gpu
TOL = 1e-3


def test_predcls_eval_forward_parity(cuda):
    """BASELINE config 1: PredCls forward, 1 synthetic 592x592 image, 20 GT boxes -> 380 pairs."""
    prod, orc = build_pair('predcls', seed=0)
    nb = make_numpy_batch(1, seed=0)
    prod = prod.to(cuda).eval(); orc.eval()
    prod.keep_last_result = True
    with torch.no_grad():
        pb, po, ps, pr, pp = prod(*to_tuple(nb, cuda))
        ob, oo, os_, or_, op = orc(torch.from_numpy(nb["imgs"]), nb["im_sizes"], 0, torch.from_numpy(nb["gt_boxes"]),
                                   torch.from_numpy(nb["gt_classes"]), torch.from_numpy(nb["gt_rels"]))
    assert np.array_equal(pb, ob) and np.array_equal(po, oo)
    np.testing.assert_allclose(ps, os_, rtol=1e-5, atol=1e-6)
    assert pr.shape == (380, 2) and pp.shape == (380, 51)
    # same set of pairs; predicate distributions compared pair by pair (ordering of near-ties may differ)
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    ip, io = np.argsort(key(pr)), np.argsort(key(or_))
    assert np.array_equal(pr[ip], or_[io])
    # the bar is on the fp32 LOGITS (pred_scores are their softmax, which turns a 1e-3 relative logit
    # error into a several-percent probability error when |logit| is in the hundreds, as it is here)
    lp, lo = prod.last_result.rel_dists.cpu(), orc.last_result.rel_dists
    assert relerr(lp, lo) < TOL, (relerr(lp, lo), float(lo.abs().max()))
    scale = float(lo.abs().max())
    big = op[io] > 1e-4
    assert np.abs(np.log(pp[ip][big]) - np.log(op[io][big])).max() < 2 * TOL * scale
    # and the product's own order is sorted by its triple score (surgery.py:47-49)
    sc = pp[:, 1:].max(1) * ps[pr[:, 0]] * ps[pr[:, 1]]
    assert (np.diff(sc) <= 1e-7).all()


@pytest.mark.parametrize("B,boxes", [(2, 9), (6, 20)])
def test_sgcls_train_step_parity(cuda, B, boxes):
    """BASELINE config 2 (and a small variant): SGCls training forward + backward. Same sampled
    relation triples (shared RNG stream), logits within TOL, parameter gradients within TOL."""
    prod, orc = build_pair('sgcls', seed=1)
    nb = make_numpy_batch(B, seed=3, boxes_per_img=boxes, rels_per_img=min(15, boxes))
    prod = prod.to(cuda).train(); orc.train()
    n_obj = nb["gt_boxes"].shape[0]
    n_rel = min(B * boxes * (boxes - 1), 256 * B)
    det, top, ctx = make_masks(n_obj, n_rel, B, seed=5)
    prod.detector.dropout_masks = to_dev(det, cuda); prod.dropout_masks = to_dev(top, cuda)
    prod.context.dropout_masks = to_dev(ctx, cuda)
    orc.detector.masks = det; orc.masks = top; orc.context.masks = ctx
    prod.detector.rng = np.random.RandomState(11); orc.detector.rng = np.random.RandomState(11)

    res_p = prod(*to_tuple(nb, cuda))
    res_o = orc(torch.from_numpy(nb["imgs"]), nb["im_sizes"], 0, torch.from_numpy(nb["gt_boxes"]),
                torch.from_numpy(nb["gt_classes"]), torch.from_numpy(nb["gt_rels"]))
    assert torch.equal(res_p.rel_labels.cpu(), res_o.rel_labels)
    assert torch.equal(res_p.rm_obj_labels.cpu(), res_o.rm_obj_labels)
    assert res_p.rel_dists.shape == (n_rel, 51)
    assert relerr(res_p.rm_obj_dists.detach(), res_o.rm_obj_dists.detach()) < TOL
    assert relerr(res_p.rel_dists.detach(), res_o.rel_dists.detach()) < TOL

    F = torch.nn.functional
    loss_p = F.cross_entropy(res_p.rm_obj_dists, res_p.rm_obj_labels) + F.cross_entropy(res_p.rel_dists, res_p.rel_labels[:, -1])
    loss_o = F.cross_entropy(res_o.rm_obj_dists, res_o.rm_obj_labels) + F.cross_entropy(res_o.rel_dists, res_o.rel_labels[:, -1])
    assert abs(float(loss_p) - float(loss_o)) < TOL * max(1.0, abs(float(loss_o)))
    loss_p.backward(); loss_o.backward()
    gp = dict(prod.named_parameters()); go = dict(orc.named_parameters())
    checked = 0
    for name in ["rel_compress.weight", "rel_compress.bias", "post_lstm.weight", "context.edge_ctx_rnn.weight",
                 "context.edge_ctx_rnn.bias", "context.obj_ctx_rnn.weight", "context.decoder_rnn.out.weight",
                 "context.decoder_rnn.input_linearity.weight", "context.decoder_rnn.state_linearity.weight",
                 "context.obj_embed.weight", "context.pos_embed.1.weight", "roi_fmap.1.3.weight", "roi_fmap.1.0.bias",
                 "roi_fmap_obj.3.weight", "union_boxes.conv.4.weight", "freq_bias.obj_baseline.weight"]:
        assert gp[name].grad is not None, name
        assert l2err(gp[name].grad, go[name].grad) < 5 * TOL, (name, l2err(gp[name].grad, go[name].grad))
        checked += 1
    assert checked == 16
    assert all(p.grad is None for p in prod.detector.parameters())


def test_sgdet_eval_runs_and_is_consistent(cuda):
    """BASELINE config 3's path (RPN + NMS on, VGG backbone): the product's detections are checked for
    internal consistency; RPN proposals are compared with the oracle on IDENTICAL head outputs."""
    prod, orc = build_pair('sgdet', seed=2)
    prod = prod.to(cuda).eval(); orc.eval()
    prod.detector.thresh = orc.detector.thresh = 0.0   # random weights: scores ~ 1/151
    nb = make_numpy_batch(1, seed=7)
    with torch.no_grad():
        feats = torch.randn(1, 37, 37, 20, 6, generator=torch.Generator().manual_seed(0))
        rp = prod.detector.rpn_head.roi_proposals(feats.to(cuda), nb["im_sizes"], 0.7, 6000, 1000).cpu()
        ro = orc.detector.rpn_head.roi_proposals(feats, nb["im_sizes"], 0.7, 6000, 1000)
        assert rp.shape == ro.shape
        assert float((rp - ro).abs().max()) < 1e-2       # expf ulp differences in the decode only
        out = prod(*to_tuple(nb, cuda))
    boxes, objs, scores, rels, pred = out
    assert boxes.shape[0] == objs.shape[0] == scores.shape[0] <= 64
    assert (objs > 0).all() and rels.shape[1] == 2 and pred.shape[1] == 51
    assert (rels[:, 0] != rels[:, 1]).all()
    sc = pred[:, 1:].max(1) * scores[rels[:, 0]] * scores[rels[:, 1]]
    assert (np.diff(sc) <= 1e-7).all()


def test_sgdet_train_step_runs(cuda):
    """SGDet TRAINING forward+backward (scripts/refine_for_detection.sh): RPN -> NMS -> per-class NMS ->
    IoU relabelling -> rel_assignments -> step-loop decoder (background labels) -> losses. No oracle for
    this path (stochastic detections); checks shapes, label ranges, finite losses and gradients."""
    from lib.fpn.anchor_targets import anchor_target_layer
    prod, _ = build_pair('sgdet', seed=4)
    prod = prod.to(cuda).train()
    prod.detector.thresh = 0.0
    prod.detector.rng = np.random.RandomState(0)
    B = 2
    nb = make_numpy_batch(B, seed=8, boxes_per_img=10, rels_per_img=8)
    tai = []
    for i in range(B):
        gb = nb["gt_boxes"][nb["gt_classes"][:, 0] == i]
        _, inds, _, _ = anchor_target_layer(gb, (592, 592), rng=np.random.RandomState(i))
        tai.append(np.column_stack((np.full(inds.shape[0], i), inds)))
    tai = torch.from_numpy(np.concatenate(tai).astype(np.int64)).to(cuda)
    tup = list(to_tuple(nb, cuda)); tup[7] = tai
    res = prod(*tup)
    n = res.rm_obj_dists.size(0)
    assert res.rm_obj_labels.shape == (n,) and int(res.rm_obj_labels.min()) >= 0
    assert res.rel_labels.size(1) == 4 and res.rel_dists.shape == (res.rel_labels.size(0), 51)
    assert int(res.rel_labels[:, 1].max()) < n and int(res.rel_labels[:, 2].max()) < n
    F = torch.nn.functional
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    loss.backward()
    assert torch.isfinite(loss)
    g = prod.context.decoder_rnn.input_linearity.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
