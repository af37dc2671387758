"""Generates tests/golden/reference_sg_eval.npz by RUNNING THE REFERENCE's lib/evaluation/sg_eval.py in this
container (its bbox.pyx compiled as is in oracle/_ref, h5py stubbed) on seeded synthetic images: every mode
(predcls, sgcls, sgdet, phrdet, preddet), with and without multiple_preds, incl. an image where nothing matches.
The fixture stores inputs and outputs; tests/test_sg_eval.py replays the inputs through this repo's restatement.

    python tests/golden/make_golden_sg_eval.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (import_reference)


def softmax(x):
    e = np.exp(x - x.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


def make_case(rng, G, R, jitter, extra, mislabel):
    gt_boxes = MG.rand_boxes(rng, G, lo=30.0).astype(np.float32)
    gt_classes = rng.randint(1, 151, G).astype(np.int64)
    pairs = np.array([(i, j) for i in range(G) for j in range(G) if i != j])
    sel = pairs[rng.choice(len(pairs), R, replace=False)]
    gt_rels = np.column_stack((sel, rng.randint(1, 51, R))).astype(np.int64)
    # detections: jittered GT boxes (some beyond IoU 0.5) + spurious boxes; some labels wrong
    pb = gt_boxes + rng.uniform(-jitter, jitter, gt_boxes.shape).astype(np.float32)
    pb = np.concatenate((pb, MG.rand_boxes(rng, extra, lo=30.0)), 0).astype(np.float32)
    pc = np.concatenate((gt_classes, rng.randint(1, 151, extra)))
    wrong = rng.rand(pc.shape[0]) < mislabel
    pc = np.where(wrong, rng.randint(1, 151, pc.shape[0]), pc).astype(np.int64)
    obj_scores = rng.uniform(0.2, 1.0, pc.shape[0])
    N = pc.shape[0]
    allp = np.array([(i, j) for i in range(N) for j in range(N) if i != j])
    logits = rng.randn(len(allp), 51) * 2.0
    for r in gt_rels:                       # make a share of the GT predicates likely
        if rng.rand() < 0.7:
            row = np.flatnonzero((allp[:, 0] == r[0]) & (allp[:, 1] == r[1]))[0]
            logits[row, r[2]] += 6.0
    rel_scores = softmax(logits)
    # the model's filter_dets order: by obj_score(subj) * obj_score(obj) * best non-background predicate score
    key = obj_scores[allp[:, 0]] * obj_scores[allp[:, 1]] * rel_scores[:, 1:].max(1)
    order = np.argsort(-key)
    return dict(gt_boxes=gt_boxes, gt_classes=gt_classes, gt_rels=gt_rels, pred_boxes=pb, pred_classes=pc,
                obj_scores=obj_scores, pred_rel_inds=allp[order], rel_scores=rel_scores[order])


def main():
    MG.import_reference()
    from lib.evaluation import sg_eval as ref      # the reference's module
    rng = np.random.RandomState(77)
    out = {}
    cases = []
    for ci, (G, R, jitter, extra, mislabel) in enumerate([(12, 8, 6.0, 4, 0.1), (20, 15, 25.0, 6, 0.3),
                                                          (5, 3, 2.0, 0, 0.0), (9, 6, 200.0, 3, 1.0)]):
        cases.append(make_case(rng, G, R, jitter, extra, mislabel))
    runs = [(m, False) for m in ('predcls', 'sgcls', 'sgdet')] + [('sgdet', True), ('phrdet', True), ('preddet', True)]
    n = 0
    for ci, c in enumerate(cases):
        for k, v in c.items():
            out["case%d_%s" % (ci, k)] = v
        G = c["gt_boxes"].shape[0]
        for mode, multi in runs:
            gt_entry = dict(gt_relations=c["gt_rels"], gt_boxes=c["gt_boxes"], gt_classes=c["gt_classes"])
            if mode in ('predcls', 'sgcls', 'preddet'):
                # GT boxes are given: only pairs among the first G boxes exist
                keep = (c["pred_rel_inds"] < G).all(1)
                pe = dict(pred_rel_inds=c["pred_rel_inds"][keep], rel_scores=c["rel_scores"][keep],
                          pred_classes=c["pred_classes"][:G], obj_scores=c["obj_scores"][:G], pred_boxes=c["gt_boxes"])
            else:
                pe = dict(pred_rel_inds=c["pred_rel_inds"], rel_scores=c["rel_scores"], pred_classes=c["pred_classes"],
                          obj_scores=c["obj_scores"], pred_boxes=c["pred_boxes"])
            ev = ref.BasicSceneGraphEvaluator(mode, multiple_preds=multi)
            res = ev.evaluate_scene_graph_entry(gt_entry, pe)
            tag = "run%d" % n
            out[tag + "_meta"] = np.array([ci, ('predcls', 'sgcls', 'sgdet', 'phrdet', 'preddet').index(mode), int(multi)])
            out[tag + "_recall"] = np.array([ev.result_dict[mode + '_recall'][k][0] for k in (20, 50, 100)])
            if res[0] is not None:
                p2g = res[0]
                out[tag + "_p2g_len"] = np.array([len(x) for x in p2g], dtype=np.int64)
                out[tag + "_p2g_val"] = np.array([g for x in p2g for g in x], dtype=np.int64)
                out[tag + "_5ples"] = np.asarray(res[1])
                out[tag + "_scores"] = np.asarray(res[2])
            n += 1
    out["n_runs"] = np.array(n)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "reference_sg_eval.npz"), **out)
    print("wrote reference_sg_eval.npz:", n, "runs; recalls:",
          [np.round(out["run%d_recall" % i], 3).tolist() for i in range(n)])


if __name__ == "__main__":
    main()
