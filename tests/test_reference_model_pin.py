"""MODEL-LEVEL pin: the oracle's RelModel against the REFERENCE's own RelModel executed on the CPU in the build
container (tests/golden/make_golden_model.py): every Python line of the reference model ran — context construction,
sorting / packing, decoder loop, union-box branch, relation tail, frequency bias, filter_dets — with its three CUDA
extensions replaced by the oracle's operator restatements (pinned separately on the GPU against the reference's .cu
files) and GloVe / VG / ImageNet tables replaced by seeded synthetic values. Both sides load the same synthetic state
dict, regenerated from (name, shape, seed) by tests/golden/synthetic_state.py; the fixture holds the reference's
state-dict keys / shapes and its eval outputs for BASELINE config 1 (one 592x592 image, 20 GT boxes, 380 pairs)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


_CACHE = {}


def script_config_oracle(mode, **attrs):
    """ONE oracle RelModel in the script configuration serves every test of this file that uses it: the state dict does
    not depend on the mode / ordering / tail flags, which are plain attributes (building the three VGG fc stacks and
    regenerating 1.7 GB of synthetic weights per test would triple the run time of the CPU suite)."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW
    if "orc" not in _CACHE:
        orc = OM.RelModel(CLASSES, RELS, mode="sgcls", **KW)
        sd = orc.state_dict()
        _CACHE["state"] = synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3)
        _CACHE["orc"] = orc
    orc = _CACHE["orc"]
    orc.load_state_dict(_CACHE["state"])                       # a training-mode test may have moved the BatchNorm buffers
    orc.mode = orc.context.mode = mode
    orc.detector.mode = 'refinerels' if mode == 'sgdet' else 'gtbox'
    orc.require_overlap = mode == 'sgdet'
    orc.detector.thresh = attrs.pop("thresh", 0.01)
    orc.context.order = attrs.pop("order", KW["order"])
    orc.use_tanh = attrs.pop("use_tanh", KW["use_tanh"])
    orc.limit_vision = attrs.pop("limit_vision", KW["limit_vision"])
    assert not attrs, attrs
    for p_ in orc.parameters():
        p_.requires_grad = True
        p_.grad = None
    orc.masks = orc.detector.masks = orc.context.masks = None
    return orc


@pytest.mark.parametrize("mode", ["predcls", "sgcls"])
def test_oracle_relmodel_eval_matches_reference_relmodel(mode):
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    orc = script_config_oracle(mode)
    sd = orc.state_dict()
    ref_keys = [str(k) for k in g[mode + "_keys"]]
    ref_shapes = {k: tuple(int(v) for v in s.split(";") if v) for k, s in zip(ref_keys, g[mode + "_shapes"])}
    # the reference's state dict and the oracle's (= the product's) are interchangeable: same keys, same shapes
    assert set(sd.keys()) == set(ref_keys), (set(sd) ^ set(ref_keys))
    assert all(tuple(sd[k].shape) == ref_shapes[k] for k in ref_keys)
    orc.eval()
    nb = make_inputs(seed=11)
    t = torch.from_numpy
    with torch.no_grad():
        boxes, objs, obj_scores, rels, pred_scores = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]),
                                                         t(nb["gt_classes"]), t(nb["gt_rels"]))
    assert np.array_equal(np.asarray(boxes), g[mode + "_boxes"])
    assert np.array_equal(np.asarray(objs), g[mode + "_objs"])
    assert np.allclose(np.asarray(obj_scores), g[mode + "_obj_scores"], rtol=1e-4, atol=1e-6)
    want_rels, want_scores = g[mode + "_rels"], g[mode + "_pred_scores"]
    assert np.asarray(rels).shape == want_rels.shape == (380, 2)
    # same relation -> same predicate distribution (order-independent), then the ranking itself
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])
    err = np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max()
    assert err < 1e-4 * max(1.0, float(np.abs(want_scores).max())), err
    assert (np.asarray(rels) == want_rels).all(1).mean() > 0.98      # ranking equal up to near-ties of the sort key


def test_oracle_relmodel_sgcls_train_forward_matches_reference_relmodel():
    """Training forward (models/train_rels.py:118-141): GT-box relation sampling, training-mode BatchNorm in the
    position embedding and the union-box branch, teacher-forced decoder, the two cross-entropies; dropout off on both
    sides (the reference draws its masks from torch's RNG inside nn.Dropout, the oracle takes injected masks)."""
    import torch.nn.functional as F
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    from model_utils import make_masks
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_train.npz"))
    orc = script_config_oracle("sgcls")
    orc.train()
    nb = make_inputs(seed=12, boxes=14, rels=9)
    n_obj, n_rel = 14, g["train_rel_labels"].shape[0]
    det, top, ctx = make_masks(n_obj, n_rel, 1, seed=0)
    ones = lambda d: {k: torch.ones_like(v) for k, v in d.items()}
    orc.detector.masks, orc.masks, orc.context.masks = ones(det), ones(top), ones(ctx)
    orc.detector.rng = np.random.RandomState(21)
    t = torch.from_numpy
    res = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    assert np.array_equal(res.rel_labels.numpy(), g["train_rel_labels"])
    assert np.array_equal(res.rm_obj_labels.numpy(), g["train_rm_obj_labels"])
    for k, tol in (("rm_obj_dists", 1e-4), ("rel_dists", 1e-4)):
        got, want = getattr(res, k).detach().numpy(), g["train_" + k]
        assert np.abs(got - want).max() < tol * max(1.0, float(np.abs(want).max())), (k, np.abs(got - want).max())
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    assert abs(float(loss) - float(g["train_loss"])) < 1e-4 * float(g["train_loss"])
    assert np.allclose(orc.union_boxes.conv[2].running_mean.numpy(), g["train_bn_running_mean"], rtol=1e-4, atol=1e-6)
    # backward: every trainable parameter of the reference model received a gradient; norms and 16 samples each agree
    for p_ in orc.detector.parameters():
        p_.requires_grad = False
    loss.backward()
    grads = {k: p_.grad for k, p_ in orc.named_parameters() if p_.grad is not None}
    names = [str(k) for k in g["train_grad_names"]]
    assert set(names) == set(grads), set(names) ^ set(grads)
    for k, norm, samp in zip(names, g["train_grad_norms"], g["train_grad_samples"]):
        gf = grads[k].reshape(-1)
        idx = (torch.arange(16) * (gf.numel() - 1)) // 15
        assert abs(float(gf.double().norm()) - norm) < 1e-3 * max(norm, 1e-8), (k, float(gf.double().norm()), norm)
        assert np.abs(gf[idx].numpy() - samp).max() < 1e-3 * max(float(np.abs(samp).max()), float(norm) / gf.numel() ** 0.5, 1e-9), k


def test_oracle_relmodel_sgdet_eval_matches_reference_relmodel():
    """SGDet eval: RPN head, proposal decode + NMS, detector heads, per-class NMS to 64 detections, overlapping pairs,
    context with the decoder's overlap-aware commitments, relation tail, filter_dets (detector threshold 0)."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    orc = script_config_oracle("sgdet", thresh=0.0)
    ref_keys = [str(k) for k in g["sgdet_keys"]]
    assert set(orc.state_dict().keys()) == set(ref_keys)
    orc.eval()
    nb = make_inputs(seed=11)
    with torch.no_grad():
        boxes, objs, obj_scores, rels, pred_scores = orc(torch.from_numpy(nb["imgs"]), nb["im_sizes"], 0)
    assert np.asarray(boxes).shape == g["sgdet_boxes"].shape == (64, 4)
    assert np.abs(np.asarray(boxes) - g["sgdet_boxes"]).max() < 1e-2            # pixels, after exp() of the regressed deltas
    assert np.array_equal(np.asarray(objs), g["sgdet_objs"])
    assert np.allclose(np.asarray(obj_scores), g["sgdet_obj_scores"], rtol=1e-3, atol=1e-6)
    want_rels, want_scores = g["sgdet_rels"], g["sgdet_pred_scores"]
    assert np.asarray(rels).shape == want_rels.shape
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])                    # the same candidate pairs
    err = np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max()
    assert err < 1e-3 * max(1.0, float(np.abs(want_scores).max())), err
    assert (np.asarray(rels) == want_rels).all(1).mean() > 0.95


def test_oracle_resnet_detector_matches_reference_detector():
    """ObjectDetector(use_resnet=True) of the reference (object_detector.py:84-138, torchvision resnet101 conv1..layer3,
    compress, SELU roi_fmap) in GT-box eval mode, run on the CPU, against the oracle's ResNet branch (row a1')."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_resnet_detector.npz"))
    orc = OM.ObjectDetector(CLASSES, mode="gtbox", use_resnet=True)
    sd = orc.state_dict()
    ref_keys = [str(k) for k in g["keys"]]
    assert set(sd.keys()) == set(ref_keys), (set(sd) ^ set(ref_keys))
    orc.load_state_dict(synthetic_state([(k, tuple(sd[k].shape), sd[k].dtype) for k in ref_keys], seed=5))
    orc.eval()
    nb = make_inputs(seed=13, boxes=12, rels=5)
    t = torch.from_numpy
    with torch.no_grad():
        r = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]))
    scale = float(g["fmap_absmax"])
    assert np.abs(r.fmap.numpy()[0, ::64, ::4, ::4] - g["fmap_sample"]).max() < 1e-4 * scale
    assert np.abs(r.obj_fmap.numpy() - g["obj_fmap"]).max() < 1e-4 * max(1.0, float(np.abs(g["obj_fmap"]).max()))
    assert np.abs(r.od_obj_dists.numpy() - g["od_obj_dists"]).max() < 1e-4 * max(1.0, float(np.abs(g["od_obj_dists"]).max()))


def test_oracle_detector_training_forward_matches_reference_detector():
    """ObjectDetector(mode='rpntrain').train() of the reference run on the CPU: RPN
    scores / deltas at the sampled anchors, 2000 proposals -> proposal_assignments_det (stable candidate order, numpy RNG)
    -> 256 rois with labels and box targets -> detection heads; the inputs of models/train_detector.py's four losses."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_detector_train.npz"))
    orc = OM.ObjectDetector(CLASSES, mode="rpntrain")
    sd = orc.state_dict()
    orc.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=6))
    orc.train()
    nb = make_inputs(seed=14, boxes=10, rels=4)
    n = g["od_obj_labels"].shape[0]
    orc.masks = {"roi_fmap.2": torch.ones(n, 4096), "roi_fmap.5": torch.ones(n, 4096)}
    orc.rng = np.random.RandomState(31)
    t = torch.from_numpy
    r = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), None, t(g["train_anchor_inds"]))
    assert np.array_equal(r.od_obj_labels.numpy(), g["od_obj_labels"]) and int((g["od_obj_labels"] > 0).sum()) > 0
    assert np.allclose(r.od_box_priors.numpy(), g["od_box_priors"], rtol=0, atol=2e-3)        # pixels
    assert np.array_equal(r.od_box_targets.numpy(), g["od_box_targets"])
    for k in ("rpn_scores", "rpn_box_deltas", "od_obj_dists", "od_box_deltas"):
        got, want = getattr(r, k).detach().numpy(), g[k]
        assert np.abs(got - want).max() < 1e-4 * max(1.0, float(np.abs(want).max())), (k, np.abs(got - want).max())


@pytest.mark.parametrize("tag,mode,extra", [("var_conf", "predcls", dict(order="confidence", use_tanh=True, limit_vision=True)),
                                            ("var_size", "sgcls", dict(order="size"))])
def test_oracle_relmodel_constructor_variants_match_reference(tag, mode, extra):
    """Object ordering by confidence / by box size (rel_model.py:139-161) and the tanh + limit_vision relation tail
    (:515-522) — constructor arguments of the reference's RelModel other than the script configuration."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    orc = script_config_oracle(mode, **extra)
    orc.eval()
    nb = make_inputs(seed=15, boxes=16, rels=6)
    t = torch.from_numpy
    with torch.no_grad():
        boxes, objs, obj_scores, rels, pred_scores = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]),
                                                         t(nb["gt_classes"]), t(nb["gt_rels"]))
    assert np.array_equal(np.asarray(objs), g[tag + "_objs"])
    assert np.allclose(np.asarray(obj_scores), g[tag + "_obj_scores"], rtol=1e-4, atol=1e-6)
    want_rels, want_scores = g[tag + "_rels"], g[tag + "_pred_scores"]
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])
    assert np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max() < 1e-4


def test_oracle_relmodel_reference_default_arguments_predcls():
    """`RelModel(classes, rel_classes, mode='predcls')` with the reference's DEFAULT constructor arguments
    (rel_model.py:303-308): hidden 256, pooling 2048 (fc7 dropped from the union branch, limit_vision slicing), nl_obj 1,
    nl_edge 2, ordering by confidence, object features passed to the edge LSTM, tanh."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    orc = OM.RelModel(CLASSES, RELS, mode="predcls", embed_dim=200, hidden_dim=256, pooling_dim=2048, nl_obj=1, nl_edge=2,
                      order="confidence", thresh=0.01, use_bias=True, use_tanh=True, limit_vision=True,
                      pass_in_obj_feats_to_decoder=True, pass_in_obj_feats_to_edge=True)
    sd = orc.state_dict()
    ref_keys = [str(k) for k in g["var_default_keys"]]
    ref_shapes = {k: tuple(int(v) for v in s.split(";") if v) for k, s in zip(ref_keys, g["var_default_shapes"])}
    assert set(sd.keys()) == set(ref_keys), (set(sd) ^ set(ref_keys))
    assert all(tuple(sd[k].shape) == ref_shapes[k] for k in ref_keys), [k for k in ref_keys if tuple(sd[k].shape) != ref_shapes[k]]
    orc.load_state_dict(synthetic_state([(k, ref_shapes[k], sd[k].dtype) for k in ref_keys], seed=3))
    orc.eval()
    nb = make_inputs(seed=16, boxes=11, rels=5)
    t = torch.from_numpy
    with torch.no_grad():
        boxes, objs, obj_scores, rels, pred_scores = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]),
                                                         t(nb["gt_classes"]), t(nb["gt_rels"]))
    want_rels, want_scores = g["var_default_rels"], g["var_default_pred_scores"]
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])
    assert np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max() < 1e-4


def test_oracle_relmodel_sgdet_train_forward_matches_reference_relmodel():
    """SGDet TRAINING forward (scripts/refine_for_detection.sh): RPN proposals -> per-class NMS detections -> IoU
    relabelling against the GT boxes -> rel_assignments on the detected boxes (numpy RNG) -> context with the decoder
    teacher-forced on labels that contain background -> both cross-entropies. The fixture's GT boxes are built from the
    model's own detections so that labelled detections and foreground relations exist."""
    import torch.nn.functional as F
    from golden.synthetic_state import make_inputs
    from model_utils import make_masks
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_train.npz"))
    orc = script_config_oracle("sgdet", thresh=0.0)
    orc.train()
    nb = make_inputs(seed=11)
    n_det, n_rel = g["sgdet_train_rm_obj_labels"].shape[0], g["sgdet_train_rel_labels"].shape[0]
    det, top, ctx = make_masks(n_det, n_rel, 1, seed=0)
    ones = lambda d: {k: torch.ones_like(v) for k, v in d.items()}
    orc.masks, orc.context.masks = ones(top), ones(ctx)
    orc.detector.masks = {"roi_fmap.2": torch.ones(1, 4096), "roi_fmap.5": torch.ones(1, 4096)}      # broadcast over the rois
    orc.detector.rng = np.random.RandomState(41)
    t = torch.from_numpy
    res = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(g["sgdet_train_gt_boxes"]), t(g["sgdet_train_gt_classes"]),
              t(g["sgdet_train_gt_rels"]))
    assert np.array_equal(res.rm_obj_labels.numpy(), g["sgdet_train_rm_obj_labels"]) and int((res.rm_obj_labels > 0).sum()) > 5
    assert np.array_equal(res.rel_labels.numpy(), g["sgdet_train_rel_labels"]) and int((res.rel_labels[:, -1] > 0).sum()) > 0
    for k, tol in (("rm_obj_dists", 1e-3), ("rel_dists", 1e-3)):
        got, want = getattr(res, k).detach().numpy(), g["sgdet_train_" + k]
        assert np.abs(got - want).max() < tol * max(1.0, float(np.abs(want).max())), (k, np.abs(got - want).max())
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    assert abs(float(loss.detach()) - float(g["sgdet_train_loss"])) < 1e-3 * float(g["sgdet_train_loss"])


@pytest.mark.parametrize("tag,mode,thresh", [("base_sgcls", "sgcls", 0.01), ("base_sgdet", "sgdet", 0.0)])
def test_oracle_baseline_configuration_matches_reference(tag, mode, thresh):
    """The scripts' "baseline" (`-nl_obj 0 -nl_edge 0`, scripts/train_models_sgcls.sh:8, eval_models_sg*.sh): linear object
    classifier instead of the context LSTMs (rel_model.py:125-126, 259-283 incl. the per-class NMS label choice of SGDet
    eval) and `post_emb` instead of `post_lstm` (:386-388, 500-503)."""
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    if "base" not in _CACHE:                                   # one build serves both modes (same state dict)
        orc = OM.RelModel(CLASSES, RELS, mode="sgcls", **dict(KW, nl_obj=0, nl_edge=0))
        sd = orc.state_dict()
        orc.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
        _CACHE["base"] = orc
    orc = _CACHE["base"]
    orc.mode = orc.context.mode = mode
    orc.detector.mode = 'refinerels' if mode == 'sgdet' else 'gtbox'
    orc.require_overlap = mode == 'sgdet'
    orc.detector.thresh = thresh
    ref_keys = [str(k) for k in g[tag + "_keys"]]
    assert set(orc.state_dict().keys()) == set(ref_keys), (set(orc.state_dict()) ^ set(ref_keys))
    orc.eval()
    nb = make_inputs(seed=17, boxes=13, rels=5)
    t = torch.from_numpy
    with torch.no_grad():
        if mode == "sgdet":
            out = orc(t(nb["imgs"]), nb["im_sizes"], 0)
        else:
            out = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    boxes, objs, obj_scores, rels, pred_scores = out
    assert np.array_equal(np.asarray(objs), g[tag + "_objs"])
    assert np.allclose(np.asarray(obj_scores), g[tag + "_obj_scores"], rtol=1e-3, atol=1e-6)
    want_rels, want_scores = g[tag + "_rels"], g[tag + "_pred_scores"]
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])
    assert np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max() < 1e-3


@pytest.mark.parametrize("tag,mode,kw", [
    ("base_sgcls", "sgcls", dict(hidden_dim=512, pooling_dim=4096, nl_obj=0, nl_edge=0, order='leftright', use_bias=True,
                                 use_tanh=False, limit_vision=False, pass_in_obj_feats_to_decoder=False,
                                 pass_in_obj_feats_to_edge=False)),
    ("var_default", "predcls", {}),
])
def test_product_state_dict_keys_equal_the_reference_models(tag, mode, kw, monkeypatch):
    """The PRODUCT's RelModel (constructed on the CPU; no kernel runs) exposes exactly the state-dict keys of the
    reference's RelModel for the scripts' baseline and the reference's default arguments (the MotifNet script
    configuration is covered by every GPU model test, which loads one state dict into oracle and product) — checkpoints (`vgrel-*.tar`, train_rels.py:75-95) interchange."""
    from lib.rel_model import RelModel
    from golden.synthetic_state import CLASSES, RELS
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    # only names and shapes matter here: skip the (QR-based, slow) orthogonal initialisation of the LSTM weights
    monkeypatch.setattr(torch.nn.init, "orthogonal_", lambda tensor, gain=1: tensor)
    prod = RelModel(CLASSES, RELS, mode=mode, **kw)
    ref_keys = [str(k) for k in g[tag + "_keys"]]
    assert set(prod.state_dict().keys()) == set(ref_keys), (set(prod.state_dict()) ^ set(ref_keys))
    if tag + "_shapes" in g:
        shapes = {k: tuple(int(v) for v in s.split(";") if v) for k, s in zip(ref_keys, g[tag + "_shapes"])}
        assert all(tuple(v.shape) == shapes[k] for k, v in prod.state_dict().items())


def test_oracle_sgdet_eval_from_precomputed_proposals_matches_reference():
    """`use_proposals=True`: the detector takes 2000 scored boxes per image instead of running the RPN
    (object_detector.py:216-258, filter_roi_proposals :600-612); everything downstream as in SGDet eval."""
    from golden.synthetic_state import make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_model_eval.npz"))
    orc = script_config_oracle("sgdet", thresh=0.0)
    orc.detector.mode = 'proposals'
    orc.eval()
    nb = make_inputs(seed=19)
    with torch.no_grad():
        boxes, objs, obj_scores, rels, pred_scores = orc(torch.from_numpy(nb["imgs"]), nb["im_sizes"], 0,
                                                         proposals=torch.from_numpy(g["prop_proposals"]))
    assert np.array_equal(np.asarray(objs), g["prop_objs"])
    assert np.abs(np.asarray(boxes) - g["prop_boxes"]).max() < 1e-2
    want_rels, want_scores = g["prop_rels"], g["prop_pred_scores"]
    key = lambda r: r[:, 0] * 1000 + r[:, 1]
    a, b = np.argsort(key(np.asarray(rels))), np.argsort(key(want_rels))
    assert np.array_equal(np.asarray(rels)[a], want_rels[b])
    assert np.abs(np.asarray(pred_scores)[a] - want_scores[b]).max() < 1e-3


def test_train_rels_checkpoint_loading_lines_run_against_the_product(monkeypatch):
    """models/train_rels.py:84-95 executed against the PRODUCT on the CPU (no kernel runs): a detector checkpoint
    (state dict of an ObjectDetector, as `vg-24.tar`) is restored into `detector.detector` with optimistic_restore and
    its fc6 / fc7 copied into `roi_fmap[1][0|3]` and `roi_fmap_obj[0|3]` — SURVEY.md section 8b/f4."""
    from lib.rel_model import RelModel
    from lib.object_detector import ObjectDetector
    from lib.pytorch_misc import optimistic_restore
    from golden.synthetic_state import CLASSES, RELS, KW
    monkeypatch.setattr(torch.nn.init, "orthogonal_", lambda tensor, gain=1: tensor)
    src = ObjectDetector(CLASSES, mode='gtbox')                   # what train_detector.py saves
    ckpt = {'state_dict': {k: v.clone() for k, v in src.state_dict().items()}, 'epoch': 24}
    for v in ckpt['state_dict'].values():
        if v.dtype.is_floating_point:
            v.normal_()
    detector = RelModel(CLASSES, RELS, mode='sgcls', num_gpus=1, pass_in_obj_feats_to_decoder=False,
                        pass_in_obj_feats_to_edge=False, **KW)
    assert optimistic_restore(detector.detector, ckpt['state_dict'])            # every key matched, train_rels.py:85
    detector.roi_fmap[1][0].weight.data.copy_(ckpt['state_dict']['roi_fmap.0.weight'])
    detector.roi_fmap[1][3].weight.data.copy_(ckpt['state_dict']['roi_fmap.3.weight'])
    detector.roi_fmap[1][0].bias.data.copy_(ckpt['state_dict']['roi_fmap.0.bias'])
    detector.roi_fmap[1][3].bias.data.copy_(ckpt['state_dict']['roi_fmap.3.bias'])
    detector.roi_fmap_obj[0].weight.data.copy_(ckpt['state_dict']['roi_fmap.0.weight'])
    detector.roi_fmap_obj[3].weight.data.copy_(ckpt['state_dict']['roi_fmap.3.weight'])
    detector.roi_fmap_obj[0].bias.data.copy_(ckpt['state_dict']['roi_fmap.0.bias'])
    detector.roi_fmap_obj[3].bias.data.copy_(ckpt['state_dict']['roi_fmap.3.bias'])
    assert torch.equal(detector.detector.roi_fmap[0].weight, ckpt['state_dict']['roi_fmap.0.weight'])
    assert torch.equal(detector.roi_fmap_obj[3].bias, ckpt['state_dict']['roi_fmap.3.bias'])
    # the lr/10 parameter group of train_rels.py:57-61 is selected by name prefix
    fc = [n for n, p in detector.named_parameters() if n.startswith('roi_fmap') and p.requires_grad]
    assert len(fc) == 8 and all(n.startswith(('roi_fmap.1.', 'roi_fmap_obj.')) for n in fc)
    # and a "vgrel" checkpoint (the whole RelModel) round-trips
    assert optimistic_restore(detector, {k: v.clone() for k, v in detector.state_dict().items()})
