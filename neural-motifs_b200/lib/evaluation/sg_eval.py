"""Scene-graph recall (R@20/50/100) — the consumer of `RelModel`'s eval output.
Same surface as the reference's lib/evaluation/sg_eval.py: `BasicSceneGraphEvaluator(mode, multiple_preds)`
with `.evaluate_scene_graph_entry(gt_entry, pred_entry)`, `.result_dict`, `.all_modes()`, `.vrd_modes()`, and the
free functions `evaluate_from_dict`, `evaluate_recall` (sg_eval.py:11-205).

Host-side numpy like the reference's (the inputs are a few hundred boxes / triples per image). Written from the
reference's behaviour, not its text: triple matching is one vectorised comparison of packed (subject class,
predicate, object class) keys, box matching one IoU matrix per image instead of one Cython call per GT relation;
IoU is the float64 "+1 pixel" formula of lib/fpn/box_intersections_cpu/bbox.pyx:15-60. Pinned against the
reference's own module run in this container: tests/golden/reference_sg_eval.npz (tests/golden/make_golden_sg_eval.py).
"""
from functools import reduce

import numpy as np

MODES = ('sgdet', 'sgcls', 'predcls')          # config.py:23 of the reference


def intersect_2d(x1, x2):
    """[m1,n], [m2,n] -> bool [m1,m2], True where the rows are equal (lib/pytorch_misc.py:233-247)."""
    if x1.shape[1] != x2.shape[1]:
        raise ValueError("Input arrays must have same #columns")
    return (x1[:, None, :] == x2[None, :, :]).all(2)


def argsort_desc(scores):
    """Indices (one row per element, one column per dim) that sort `scores` descending (pytorch_misc.py:323-330)."""
    return np.column_stack(np.unravel_index(np.argsort(-scores.ravel()), scores.shape))


def _iou_pairwise(a, b):
    """Elementwise float64 IoU of boxes a[i] and b[i] with the +1 pixel convention of bbox.pyx."""
    a = a.astype(np.float64); b = b.astype(np.float64)
    iw = np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0]) + 1.0
    ih = np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1]) + 1.0
    inter = np.where((iw > 0) & (ih > 0), iw * ih, 0.0)
    area_a = (a[:, 2] - a[:, 0] + 1.0) * (a[:, 3] - a[:, 1] + 1.0)
    area_b = (b[:, 2] - b[:, 0] + 1.0) * (b[:, 3] - b[:, 1] + 1.0)
    ua = area_a + area_b - inter
    return np.where(inter > 0, inter / ua, 0.0)


class BasicSceneGraphEvaluator(object):
    def __init__(self, mode, multiple_preds=False):
        self.result_dict = {}
        self.mode = mode
        self.result_dict[self.mode + '_recall'] = {20: [], 50: [], 100: []}
        self.multiple_preds = multiple_preds

    @classmethod
    def all_modes(cls, **kwargs):
        return {m: cls(mode=m, **kwargs) for m in MODES}

    @classmethod
    def vrd_modes(cls, **kwargs):
        return {m: cls(mode=m, multiple_preds=True, **kwargs) for m in ('preddet', 'phrdet')}

    def evaluate_scene_graph_entry(self, gt_entry, pred_scores, viz_dict=None, iou_thresh=0.5):
        return evaluate_from_dict(gt_entry, pred_scores, self.mode, self.result_dict, viz_dict=viz_dict,
                                  iou_thresh=iou_thresh, multiple_preds=self.multiple_preds)

    def save(self, fn):
        np.save(fn, self.result_dict)

    def print_stats(self):
        print('======================' + self.mode + '============================')
        for k, v in self.result_dict[self.mode + '_recall'].items():
            print('R@%i: %f' % (k, np.mean(v)))


def evaluate_from_dict(gt_entry, pred_entry, mode, result_dict, multiple_preds=False, viz_dict=None, **kwargs):
    """gt_entry: gt_relations [R,3] (subj, obj, predicate), gt_boxes [G,4], gt_classes [G];
    pred_entry: pred_rel_inds [P,2], rel_scores [P,51], (+ pred_boxes, pred_classes, obj_scores by mode).
    Appends R@K of this image to result_dict[mode + '_recall'][K] (sg_eval.py:43-122)."""
    gt_rels = gt_entry['gt_relations']
    gt_boxes = gt_entry['gt_boxes'].astype(float)
    gt_classes = gt_entry['gt_classes']
    pred_rel_inds = pred_entry['pred_rel_inds']
    rel_scores = pred_entry['rel_scores']
    recalls = result_dict[mode + '_recall']

    if mode == 'predcls':
        pred_boxes, pred_classes = gt_boxes, gt_classes
        obj_scores = np.ones(gt_classes.shape[0])
    elif mode == 'sgcls':
        pred_boxes, pred_classes, obj_scores = gt_boxes, pred_entry['pred_classes'], pred_entry['obj_scores']
    elif mode in ('sgdet', 'phrdet'):
        pred_boxes = pred_entry['pred_boxes'].astype(float)
        pred_classes, obj_scores = pred_entry['pred_classes'], pred_entry['obj_scores']
    elif mode == 'preddet':
        # predicate detection: only the predicted pairs that are GT pairs count, ranked by predicate score
        hit = intersect_2d(pred_rel_inds, gt_rels[:, :2])
        if hit.size == 0:
            for k in recalls:
                recalls[k].append(0.0)
            return None, None, None
        first = hit.argmax(0)                                     # first predicted row equal to each GT pair
        pairs, scores = pred_rel_inds[first], rel_scores[first]
        order = argsort_desc(scores[:, 1:])
        ranked = np.column_stack((pairs[order[:, 0]], order[:, 1] + 1))
        matches = intersect_2d(ranked, gt_rels)
        for k in recalls:
            recalls[k].append(float(matches[:k].any(0).sum()) / float(gt_rels.shape[0]))
        return None, None, None
    else:
        raise ValueError('invalid mode')

    if multiple_preds:                                            # every (pair, predicate) competes; top 100 kept
        pair_score = obj_scores[pred_rel_inds].prod(1)
        order = argsort_desc(pair_score[:, None] * rel_scores[:, 1:])[:100]
        pred_rels = np.column_stack((pred_rel_inds[order[:, 0]], order[:, 1] + 1))
        predicate_scores = rel_scores[order[:, 0], order[:, 1] + 1]
    else:                                                         # one predicate per pair: the best non-background
        pred_rels = np.column_stack((pred_rel_inds, 1 + rel_scores[:, 1:].argmax(1)))
        predicate_scores = rel_scores[:, 1:].max(1)

    pred_to_gt, pred_5ples, triple_scores = evaluate_recall(
        gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes, predicate_scores, obj_scores,
        phrdet=mode == 'phrdet', **kwargs)
    for k in recalls:
        matched = reduce(np.union1d, pred_to_gt[:k])
        recalls[k].append(float(len(matched)) / float(gt_rels.shape[0]))
    return pred_to_gt, pred_5ples, triple_scores


def evaluate_recall(gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes, rel_scores=None,
                    cls_scores=None, iou_thresh=0.5, phrdet=False):
    """For every predicted triple (assumed sorted by score) the list of GT relations it matches: same
    (subject class, predicate, object class) and both boxes with IoU >= thresh (phrdet: the union boxes).
    Returns (pred_to_gt, pred_5ples [P,5] = (subj, obj, subj class, obj class, predicate), triple scores [P,3])
    (sg_eval.py:150-205)."""
    if pred_rels.size == 0:
        return [[]], np.zeros((0, 5)), np.zeros(0)
    assert gt_rels.shape[0] != 0
    assert pred_rels[:, :2].max() < pred_classes.shape[0]
    assert np.all(pred_rels[:, 2] > 0)

    gt_trip = np.column_stack((gt_classes[gt_rels[:, 0]], gt_rels[:, 2], gt_classes[gt_rels[:, 1]]))
    gt_tb = np.column_stack((gt_boxes[gt_rels[:, 0]], gt_boxes[gt_rels[:, 1]]))
    pr_trip = np.column_stack((pred_classes[pred_rels[:, 0]], pred_rels[:, 2], pred_classes[pred_rels[:, 1]]))
    pr_tb = np.column_stack((pred_boxes[pred_rels[:, 0]], pred_boxes[pred_rels[:, 1]]))
    triple_scores = None
    if rel_scores is not None and cls_scores is not None:
        triple_scores = np.column_stack((cls_scores[pred_rels[:, 0]], cls_scores[pred_rels[:, 1]], rel_scores))
        overall = triple_scores.prod(1)
        if not np.all(overall[1:] <= overall[:-1] + 1e-5):
            print("Somehow the relations weren't sorted properly: \n{}".format(overall))

    pred_to_gt = _match_triples(gt_trip, pr_trip, gt_tb, pr_tb, iou_thresh, phrdet)
    pred_5ples = np.column_stack((pred_rels[:, :2], pr_trip[:, [0, 2, 1]]))
    return pred_to_gt, pred_5ples, triple_scores


def _union(tb):
    return np.column_stack((np.minimum(tb[:, 0], tb[:, 4]), np.minimum(tb[:, 1], tb[:, 5]),
                            np.maximum(tb[:, 2], tb[:, 6]), np.maximum(tb[:, 3], tb[:, 7])))


def _match_triples(gt_trip, pr_trip, gt_tb, pr_tb, iou_thresh, phrdet):
    """pred_to_gt[p] = ascending list of GT relation indices matched by predicted triple p."""
    same = intersect_2d(gt_trip, pr_trip)                          # [R, P]
    gi, pi = np.nonzero(same)
    pred_to_gt = [[] for _ in range(pr_tb.shape[0])]
    if gi.size == 0:
        return pred_to_gt
    if phrdet:
        ok = _iou_pairwise(_union(gt_tb[gi]), _union(pr_tb[pi])) >= iou_thresh
    else:
        ok = (_iou_pairwise(gt_tb[gi, :4], pr_tb[pi, :4]) >= iou_thresh) & \
             (_iou_pairwise(gt_tb[gi, 4:], pr_tb[pi, 4:]) >= iou_thresh)
    for g, p in zip(gi[ok], pi[ok]):                               # np.nonzero is row-major: g ascending per p
        pred_to_gt[p].append(int(g))
    return pred_to_gt
