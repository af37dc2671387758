"""lib/evaluation/sg_eval.py against the REFERENCE's own lib/evaluation/sg_eval.py run in the
build container on seeded synthetic images (tests/golden/make_golden_sg_eval.py -> reference_sg_eval.npz): every mode,
with and without multiple_preds; recalls, per-prediction GT matches, 5-tuples and triple scores must be identical."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
MODES = ('predcls', 'sgcls', 'sgdet', 'phrdet', 'preddet')


def test_sg_eval_matches_reference_fixture():
    from lib.evaluation import sg_eval
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_sg_eval.npz"))
    n = int(g["n_runs"])
    assert n == 24
    seen = set()
    for i in range(n):
        tag = "run%d" % i
        ci, mi, multi = [int(v) for v in g[tag + "_meta"]]
        mode = MODES[mi]
        seen.add((mode, multi))
        c = {k: g["case%d_%s" % (ci, k)] for k in ("gt_boxes", "gt_classes", "gt_rels", "pred_boxes", "pred_classes",
                                                   "obj_scores", "pred_rel_inds", "rel_scores")}
        G = c["gt_boxes"].shape[0]
        gt_entry = dict(gt_relations=c["gt_rels"], gt_boxes=c["gt_boxes"], gt_classes=c["gt_classes"])
        if mode in ('predcls', 'sgcls', 'preddet'):
            keep = (c["pred_rel_inds"] < G).all(1)
            pe = dict(pred_rel_inds=c["pred_rel_inds"][keep], rel_scores=c["rel_scores"][keep],
                      pred_classes=c["pred_classes"][:G], obj_scores=c["obj_scores"][:G], pred_boxes=c["gt_boxes"])
        else:
            pe = dict(pred_rel_inds=c["pred_rel_inds"], rel_scores=c["rel_scores"], pred_classes=c["pred_classes"],
                      obj_scores=c["obj_scores"], pred_boxes=c["pred_boxes"])
        ev = sg_eval.BasicSceneGraphEvaluator(mode, multiple_preds=bool(multi))
        with contextlib.redirect_stdout(io.StringIO()):           # the "weren't sorted" notice of predcls, as the reference
            res = ev.evaluate_scene_graph_entry(gt_entry, pe)
        rec = np.array([ev.result_dict[mode + '_recall'][k][0] for k in (20, 50, 100)])
        assert np.array_equal(rec, g[tag + "_recall"]), (i, mode, multi, rec, g[tag + "_recall"])
        if tag + "_p2g_len" in g:
            p2g = res[0]
            assert np.array_equal(np.array([len(x) for x in p2g]), g[tag + "_p2g_len"]), (i, mode)
            assert np.array_equal(np.array([v for x in p2g for v in x], dtype=np.int64), g[tag + "_p2g_val"]), (i, mode)
            assert np.array_equal(np.asarray(res[1]), g[tag + "_5ples"])
            assert np.allclose(np.asarray(res[2]), g[tag + "_scores"], rtol=0, atol=0)
        else:
            assert res == (None, None, None)
    assert seen == {('predcls', 0), ('sgcls', 0), ('sgdet', 0), ('sgdet', 1), ('phrdet', 1), ('preddet', 1)}


def test_sg_eval_edge_cases():
    from lib.evaluation import sg_eval
    # no predictions at all: recall 0, the reference's ([[]], empty, empty) return
    gt = dict(gt_relations=np.array([[0, 1, 3]]), gt_boxes=np.array([[0, 0, 10, 10], [5, 5, 20, 20]], dtype=np.float32),
              gt_classes=np.array([4, 9]))
    pe = dict(pred_rel_inds=np.zeros((0, 2), dtype=np.int64), rel_scores=np.zeros((0, 51)))
    ev = sg_eval.BasicSceneGraphEvaluator('predcls')
    p2g, five, sc = ev.evaluate_scene_graph_entry(gt, pe)
    assert p2g == [[]] and five.shape == (0, 5) and ev.result_dict['predcls_recall'][20] == [0.0]
    # a perfect single prediction
    rs = np.zeros((2, 51)); rs[0, 3] = 0.9; rs[1, 7] = 0.8
    pe = dict(pred_rel_inds=np.array([[0, 1], [1, 0]]), rel_scores=rs)
    ev = sg_eval.BasicSceneGraphEvaluator('predcls')
    p2g, five, sc = ev.evaluate_scene_graph_entry(gt, pe)
    assert p2g == [[0], []] and ev.result_dict['predcls_recall'][100] == [1.0]
    assert five.tolist() == [[0, 1, 4, 9, 3], [1, 0, 9, 4, 7]]
    assert set(sg_eval.BasicSceneGraphEvaluator.all_modes()) == {'sgdet', 'sgcls', 'predcls'}
    assert all(e.multiple_preds for e in sg_eval.BasicSceneGraphEvaluator.vrd_modes().values())
