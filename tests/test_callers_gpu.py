"""The caller sequence of models/train_rels.py:118-152 (train_batch) and models/eval_rels.py:59-86 (val_batch) run end to
end against the PRODUCT on the GPU — restated in this repository's words (bench.train_step is the training half: the
reference source cannot travel to the GPU box; tests/test_reference_callers.py executes the reference's own function
definitions against the product's API surface in the build container).
  train: result = detector[blob]; two cross-entropies; optimizer.zero_grad(); loss.backward(); clip 5; optimizer.step()
         with the fused FlatSGD standing in for clip_grad_norm + optim.SGD — the first loss must equal the oracle's.
  eval:  det_res = detector[blob] -> (boxes, objs, obj_scores, rels, pred_scores) -> BasicSceneGraphEvaluator."""
import numpy as np
import pytest
import torch

from tests.model_utils import build_pair, make_masks, to_dev

pytestmark = pytest.mark.gpu


def test_train_batch_and_val_batch_sequences(cuda):
    import bench
    from torch.nn import functional as F
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    from config import BOX_SCALE, IM_SCALE
    from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
    from lib.evaluation.sg_eval import BasicSceneGraphEvaluator
    from lib import fused_optim
    B, boxes = 2, 8
    prod, orc = build_pair('sgcls', seed=2)
    prod = prod.to(cuda).train(); orc.train()
    nb = make_numpy_batch(B, seed=4, boxes_per_img=boxes, rels_per_img=6)
    det, top, ctx = make_masks(B * boxes, B * boxes * (boxes - 1), B, seed=6)
    prod.detector.dropout_masks = to_dev(det, cuda); prod.dropout_masks = to_dev(top, cuda); prod.context.dropout_masks = to_dev(ctx, cuda)
    orc.detector.masks, orc.masks, orc.context.masks = det, top, ctx
    prod.detector.rng = np.random.RandomState(13); orc.detector.rng = np.random.RandomState(13)
    optimizer = bench.get_optim(prod, lr=6e-3)                                # train_rels.py:57-70 (SGD branch), fused + deferred
    scheduler = ReduceLROnPlateau(optimizer, 'max', patience=3, factor=0.1, threshold=0.0001, threshold_mode='abs', cooldown=1)
    before = {n: p.detach().clone() for n, p in prod.named_parameters() if p.requires_grad}
    blob = SyntheticBlob(nb, cuda)
    loss0 = bench.train_step(prod, optimizer, blob=blob)                      # detector[b] ... optimizer.step()
    t = torch.from_numpy
    out = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    want = float(F.cross_entropy(out.rm_obj_dists, out.rm_obj_labels) + F.cross_entropy(out.rel_dists, out.rel_labels[:, -1]))
    assert abs(loss0 - want) < 1e-3 * abs(want), (loss0, want)
    prod.detector.rng = np.random.RandomState(14)
    loss1 = bench.train_step(prod, optimizer, blob=blob)
    fused_optim.wait_pending_updates(); torch.cuda.synchronize()
    assert np.isfinite(loss1)
    moved = [n for n, p in prod.named_parameters() if p.requires_grad and not torch.equal(p, before[n])]
    assert len(moved) >= 20 and all(not n.startswith("detector.") for n in moved)
    scheduler.step(0.2)                                                       # train_rels.py:199
    assert optimizer.param_groups[0]['lr'] == pytest.approx(6e-4) and optimizer.param_groups[1]['lr'] == pytest.approx(6e-3)

    # ---- eval_rels.py:59-86
    prod.eval()
    prod.detector.dropout_masks = prod.dropout_masks = prod.context.dropout_masks = None
    evaluator = BasicSceneGraphEvaluator.all_modes()
    nb1 = make_numpy_batch(1, seed=9, boxes_per_img=7, rels_per_img=6)
    with torch.no_grad():
        det_res = [prod[SyntheticBlob(nb1, cuda)]]                            # num_gpus == 1: one tuple
    for i, (boxes_i, objs_i, obj_scores_i, rels_i, pred_scores_i) in enumerate(det_res):
        gt_entry = {'gt_classes': nb1["gt_classes"][:, 1].copy(), 'gt_relations': nb1["gt_rels"][:, 1:].copy(),
                    'gt_boxes': nb1["gt_boxes"] * BOX_SCALE / IM_SCALE}
        assert np.all(objs_i[rels_i[:, 0]] > 0) and np.all(objs_i[rels_i[:, 1]] > 0)
        pred_entry = {'pred_boxes': boxes_i * BOX_SCALE / IM_SCALE, 'pred_classes': objs_i, 'pred_rel_inds': rels_i,
                      'obj_scores': obj_scores_i, 'rel_scores': pred_scores_i}
        evaluator['sgcls'].evaluate_scene_graph_entry(gt_entry, pred_entry)
    rec = evaluator['sgcls'].result_dict['sgcls_recall']
    assert all(len(rec[k]) == 1 and 0.0 <= rec[k][0] <= 1.0 for k in (20, 50, 100))
    assert boxes_i.shape == (7, 4) and rels_i.shape == (42, 2) and pred_scores_i.shape == (42, 51)
