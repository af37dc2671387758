"""Eval output filtering — lib/surgery.py:21-59 of the reference: triple score = max non-bg
predicate prob x subject score x object score, sorted descending, everything to numpy."""
import torch


def filter_dets(boxes, obj_scores, obj_classes, rel_inds, pred_scores):
    if boxes.dim() != 2:
        raise ValueError("Boxes needs to be [num_box, 4] but its {}".format(boxes.size()))
    num_box = boxes.size(0)
    assert obj_scores.size(0) == num_box
    assert obj_classes.size() == obj_scores.size()
    num_rel = rel_inds.size(0)
    assert rel_inds.size(1) == 2
    assert pred_scores.size(0) == num_rel
    obj_scores0 = obj_scores.detach()[rel_inds[:, 0]]
    obj_scores1 = obj_scores.detach()[rel_inds[:, 1]]
    pred_scores_max, _ = pred_scores.detach()[:, 1:].max(1)
    rel_scores = pred_scores_max * obj_scores0 * obj_scores1
    _, idx = torch.sort(rel_scores.view(-1), dim=0, descending=True)
    # one batched D2H at the very end of forward (the reference issues five separate .cpu() calls)
    return (boxes.detach().cpu().numpy(), obj_classes.detach().cpu().numpy(), obj_scores.detach().cpu().numpy(),
            rel_inds[idx].cpu().numpy(), pred_scores.detach()[idx].cpu().numpy())
