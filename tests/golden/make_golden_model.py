"""Generates tests/golden/reference_model_*.npz by RUNNING THE REFERENCE's own `RelModel` (lib/rel_model.py,
lib/object_detector.py, lib/get_union_boxes.py, lib/lstm/decoder_rnn.py, lib/sparse_targets.py, lib/surgery.py ...)
on the CPU in this container — the model-level orchestration that SURVEY.md §8c had to leave "unpinned".

What is the reference's own code here: every Python line of the model (context construction, sorting / packing,
decoder loop, union-box branch, relation tail, frequency bias, filter_dets, proposal sampling).
What is substituted (the reference cannot run these without PyTorch 0.3 + its CUDA extensions):
  * its three torch.utils.ffi CUDA extensions — RoIAlign, NMS, highway LSTM — are replaced by the oracle's operator
    restatements (oracle/ops.py, oracle/highway_lstm.py), which are pinned separately, on the GPU, against the
    reference's .cu files compiled unmodified (tests/test_ops_gpu.py);
  * data-dependent tables (GloVe vectors, VG frequency counts, ImageNet weights) are seeded synthetic values;
  * environment shims restore PyTorch-0.3 semantics (`Tensor.cuda` = identity, `Tensor.new(0-dim tensor)`, ...), and
    rel_assignments.py is exec'd with `async=` renamed (it does not parse on Python >= 3.7).
No reference source is edited or copied; the fixture holds inputs, the state dict's key list and the outputs.

    python tests/golden/make_golden_model.py
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import make_golden_host2 as H2  # noqa: E402


def install_shims():
    MG.import_reference()
    import itertools
    import torch
    import torch.nn as nn
    from torch.nn.utils.rnn import PackedSequence
    from oracle import ops as O
    from oracle import model as OM
    from oracle.highway_lstm import highway_lstm_forward

    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    _new = torch.Tensor.new

    def new03(self, *a, **k):          # 0.3: tensor[-1] + 1 was a Python number
        a = tuple(int(x) if isinstance(x, torch.Tensor) and x.dim() == 0 else x for x in a)
        return _new(self, *a, **k)
    torch.Tensor.new = new03
    if not hasattr(torch.nn.init, "orthogonal"):
        torch.nn.init.orthogonal = torch.nn.init.orthogonal_
    import torchvision.models.resnet as tvr
    if not hasattr(tvr, "model_urls"):
        tvr.model_urls = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    # ---- data-dependent tables -> seeded synthetic values
    gen = torch.Generator().manual_seed(1234)
    mod("lib.word_vectors", obj_edge_vectors=lambda names, wv_dim=300, **kw: torch.randn(len(names), wv_dim, generator=gen))

    def get_counts(train_data=None, must_overlap=True):
        rng = np.random.RandomState(7)
        fg = rng.randint(0, 50, (151, 151, 51)).astype(np.int64)
        bg = rng.randint(0, 500, (151, 151)).astype(np.int64)
        return fg, bg
    mod("lib.get_dataset_counts", get_counts=get_counts)

    # ---- the three CUDA extensions -> the oracle's operator restatements
    class RoIAlignFunction(object):
        def __init__(self, aligned_height, aligned_width, spatial_scale):
            self.h, self.w, self.scale = int(aligned_height), int(aligned_width), float(spatial_scale)

        def __call__(self, features, rois):
            f = features.detach().numpy()
            r = O.normalize_rois(rois.detach().numpy(), f.shape[2], f.shape[3], self.scale)
            return torch.from_numpy(O.roi_align_forward(f, r, self.h, self.w))
    mod("lib.fpn.roi_align"); mod("lib.fpn.roi_align.functions")
    mod("lib.fpn.roi_align.functions.roi_align", RoIAlignFunction=RoIAlignFunction)

    def apply_nms(scores, boxes, pre_nms_topn=12000, post_nms_topn=2000, boxes_per_im=None, nms_thresh=0.7):
        out = O.apply_nms(scores.detach().numpy(), boxes.detach().numpy(), pre_nms_topn, post_nms_topn, boxes_per_im, nms_thresh)
        if boxes_per_im is None:
            return torch.from_numpy(np.asarray(out, dtype=np.int64))
        return torch.from_numpy(np.asarray(out[0], dtype=np.int64)), out[1]
    mod("lib.fpn.nms"); mod("lib.fpn.nms.functions")
    mod("lib.fpn.nms.functions.nms", apply_nms=apply_nms)

    ns = dict(torch=torch, itertools=itertools, Variable=type("Variable03", (), {}))
    exec(H2.extract_function(os.path.join(MG.REF, "lib", "lstm", "highway_lstm_cuda", "alternating_highway_lstm.py"),
                             "block_orthogonal"), ns)

    class AlternatingHighwayLSTM(nn.Module):
        """Parameters and call convention of the reference's wrapper (alternating_highway_lstm.py:165-303); the math is
        the oracle's restatement of its CUDA kernel."""

        def __init__(self, input_size, hidden_size, num_layers=1, recurrent_dropout_probability=0):
            super().__init__()
            self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
            self.recurrent_dropout_probability = recurrent_dropout_probability
            n = sum(6 * hidden_size * (input_size if l == 0 else hidden_size) + 5 * hidden_size * hidden_size
                    for l in range(num_layers))
            self.weight = nn.Parameter(torch.randn(n) * 0.05)
            self.bias = nn.Parameter(torch.randn(5 * hidden_size * num_layers) * 0.05)

        def forward(self, inputs, initial_state=None):
            data, bs = inputs
            bs = np.asarray([int(b) for b in bs])
            T, B = len(bs), int(bs[0])
            lengths = [int((bs > b).sum()) for b in range(B)]
            off = np.concatenate(([0], np.cumsum(bs)[:-1]))
            idx_t = np.repeat(np.arange(T), bs)
            flat = torch.as_tensor(idx_t * B + (np.arange(int(bs.sum())) - off[idx_t]))
            padded = data.new_zeros(T * B, data.size(1)).index_copy(0, flat, data).view(T, B, -1)
            assert not (self.training and self.recurrent_dropout_probability > 0), "fixtures use no recurrent dropout"
            drop = torch.ones(self.num_layers, B, self.hidden_size)
            out = highway_lstm_forward(padded, lengths, self.weight, self.bias, drop, self.hidden_size, self.num_layers)
            return PS03(out.reshape(T * B, -1)[flat], [int(b) for b in bs]), None

    mod("lib.lstm.highway_lstm_cuda")
    mod("lib.lstm.highway_lstm_cuda.alternating_highway_lstm", AlternatingHighwayLSTM=AlternatingHighwayLSTM,
        block_orthogonal=ns["block_orthogonal"])

    # ---- rel_assignments.py (does not parse on Python >= 3.7)
    src = open(os.path.join(MG.REF, "lib", "fpn", "proposal_assignments", "rel_assignments.py")).read()
    ra = mod("lib.fpn.proposal_assignments.rel_assignments")
    exec(compile(src.replace("async=True", "non_blocking=True"), "rel_assignments.py", "exec"), ra.__dict__)
    return PS03


class _PS03Meta(type):
    pass


def make_ps03():
    from torch.nn.utils.rnn import PackedSequence

    class PS03(PackedSequence):
        """PyTorch-0.3 PackedSequence: a (data, batch_sizes) pair, batch_sizes a Python list."""
        def __new__(cls, data, batch_sizes, *a):
            return tuple.__new__(cls, (data, [int(b) for b in batch_sizes], None, None))

        def __iter__(self):
            return iter((tuple.__getitem__(self, 0), tuple.__getitem__(self, 1)))

        def __getitem__(self, i):
            return tuple.__getitem__(self, i)
    return PS03


PS03 = None


def main():
    global PS03
    PS03 = make_ps03()
    install_shims()
    import torch
    import torch.nn.utils.rnn as rnn_utils
    rnn_utils.PackedSequence = PS03                      # what `from torch.nn.utils.rnn import PackedSequence` now yields
    import torchvision
    import lib.object_detector as ref_od
    ref_od.vgg16 = lambda pretrained=False: torchvision.models.vgg16(weights=None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from lib.rel_model import RelModel
    print("reference RelModel imported")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
    out = {}
    for mode in ("predcls", "sgcls"):
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False,
                         use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                         rec_dropout=0.1, **KW)
        sd = m.state_dict()
        out[mode + "_keys"] = np.array(list(sd.keys()))
        out[mode + "_shapes"] = np.array([";".join(map(str, v.shape)) for v in sd.values()])
        m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
        m.eval()
        nb = make_inputs(seed=11)
        t = torch.from_numpy
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
        boxes, objs, obj_scores, rels, pred_scores = res
        for k, v in dict(boxes=boxes, objs=objs, obj_scores=obj_scores, rels=rels, pred_scores=pred_scores).items():
            out["%s_%s" % (mode, k)] = np.asarray(v)
        print(mode, "ok:", {k: np.asarray(v).shape for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res)})
    # ---- constructor variants: object ordering by confidence / size, tanh + limit_vision relation tail
    for tag, mode, kw in (("var_conf", "predcls", dict(KW, order="confidence", use_tanh=True, limit_vision=True)),
                          ("var_size", "sgcls", dict(KW, order="size"))):
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False,
                         use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                         rec_dropout=0.1, **kw)
        sd = m.state_dict()
        m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
        m.eval()
        nb = make_inputs(seed=15, boxes=16, rels=6)
        t = torch.from_numpy
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
        for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res):
            out["%s_%s" % (tag, k)] = np.asarray(v)
        print(tag, "ok")

    # ---- the reference's DEFAULT constructor arguments (rel_model.py:303-308: hidden 256, pooling 2048, nl_obj 1, nl_edge 2,
    # order confidence, pass_in_obj_feats_to_decoder / _to_edge True, tanh, limit_vision), PredCls
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = RelModel(CLASSES, RELS, mode="predcls")
    sd = m.state_dict()
    out["var_default_keys"] = np.array(list(sd.keys()))
    out["var_default_shapes"] = np.array([";".join(map(str, v.shape)) for v in sd.values()])
    m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
    m.eval()
    nb = make_inputs(seed=16, boxes=11, rels=5)
    t = torch.from_numpy
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res):
        out["var_default_" + k] = np.asarray(v)
    print("var_default ok")

    # ---- the scripts' "baseline" configuration (train_models_sgcls.sh:8, eval_models_sg*.sh: -nl_obj 0 -nl_edge 0): linear
    # object classifier instead of the context LSTMs, post_emb instead of post_lstm; SGCls and SGDet (per-class NMS labels)
    KW0 = dict(KW, nl_obj=0, nl_edge=0)
    for tag, mode, th in (("base_sgcls", "sgcls", 0.01), ("base_sgdet", "sgdet", 0.0)):
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False,
                         use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                         rec_dropout=0.1, thresh=th, **KW0)
        sd = m.state_dict()
        out[tag + "_keys"] = np.array(list(sd.keys()))
        m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
        m.eval()
        nb = make_inputs(seed=17, boxes=13, rels=5)
        t = torch.from_numpy
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if mode == "sgdet":
                res = m(t(nb["imgs"]), nb["im_sizes"], 0)
            else:
                res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
        for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res):
            out["%s_%s" % (tag, k)] = np.asarray(v)
        print(tag, "ok", np.asarray(res[3]).shape)

    # ---- SGDet eval from PRE-COMPUTED proposals (use_proposals=True -> detector mode 'proposals', object_detector.py:216-258):
    # 2000 scored boxes per image instead of the RPN
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = RelModel(CLASSES, RELS, mode="sgdet", num_gpus=1, require_overlap_det=True, use_resnet=False,
                     use_proposals=True, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                     rec_dropout=0.1, thresh=0.0, **KW)
    sd = m.state_dict()
    m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
    m.eval()
    nb = make_inputs(seed=19)
    rngp = np.random.RandomState(3)
    pb = np.concatenate([np.clip(nb["gt_boxes"] + rngp.uniform(-s_, s_, nb["gt_boxes"].shape), 0, 591) for s_ in (3, 8, 20, 40)]
                        + [MG.rand_boxes(rngp, 2000 - 4 * nb["gt_boxes"].shape[0], lo=20.0)], 0).astype(np.float32)
    props = np.column_stack((np.zeros(2000, np.float32), rngp.uniform(0.01, 1.0, 2000).astype(np.float32), pb)).astype(np.float32)
    t = torch.from_numpy
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m(t(nb["imgs"]), nb["im_sizes"], 0, None, None, None, t(props))
    out["prop_proposals"] = props
    for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res):
        out["prop_" + k] = np.asarray(v)
    print("proposals-mode sgdet ok", np.asarray(res[3]).shape)

    # ---- SGDet eval: RPN head -> proposals -> NMS -> detector -> per-class NMS -> overlapping pairs -> context with the
    # decoder's overlap-aware commitments -> relation tail (detector threshold 0 so that random weights yield detections)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = RelModel(CLASSES, RELS, mode="sgdet", num_gpus=1, require_overlap_det=True, use_resnet=False,
                     use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                     rec_dropout=0.1, thresh=0.0, **KW)
    sd = m.state_dict()
    out["sgdet_keys"] = np.array(list(sd.keys()))
    out["sgdet_shapes"] = np.array([";".join(map(str, v.shape)) for v in sd.values()])
    m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
    m.eval()
    nb = make_inputs(seed=11)
    t = torch.from_numpy
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m(t(nb["imgs"]), nb["im_sizes"], 0)
    for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res):
        out["sgdet_" + k] = np.asarray(v)
    print("sgdet ok:", {k: np.asarray(v).shape for k, v in zip("boxes objs obj_scores rels pred_scores".split(), res)})
    np.savez_compressed(os.path.join(HERE, "reference_model_eval.npz"), **out)

    # ---- the ResNet-101 detector (use_resnet=True, "Deprecated" in the reference but BASELINE config 3): GT-box mode, eval
    ref_od.resnet101 = lambda pretrained=False: torchvision.models.resnet101(weights=None)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        det = ref_od.ObjectDetector(CLASSES, mode="gtbox", use_resnet=True)
    sd = det.state_dict()
    rout = {"keys": np.array(list(sd.keys())), "shapes": np.array([";".join(map(str, v.shape)) for v in sd.values()])}
    det.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=5))
    det.eval()
    nb = make_inputs(seed=13, boxes=12, rels=5)
    t = torch.from_numpy
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = det(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), return_fmap=True)
    rout["fmap_sample"] = r.fmap.numpy()[0, ::64, ::4, ::4].copy()
    rout["fmap_absmax"] = np.array(float(r.fmap.abs().max()))
    rout["od_obj_dists"] = r.od_obj_dists.numpy()
    rout["obj_fmap"] = r.obj_fmap.numpy()
    print("resnet detector ok:", r.fmap.shape, r.od_obj_dists.shape, "fmap absmax %.3f" % float(r.fmap.abs().max()))
    np.savez_compressed(os.path.join(HERE, "reference_resnet_detector.npz"), **rout)

    # ---- detector TRAINING forward (models/train_detector.py:78-117): RPN outputs at the sampled
    # anchors, 2000 proposals, proposal -> GT assignment, detection heads. proposal_assignments_det orders its candidates
    # with torch.sort, whose order among equal keys is implementation-defined: this process runs it STABLE (as the oracle
    # and the product do), so the numpy RNG picks the same rows on both sides.
    from oracle import host as OH
    _sort = torch.sort

    def stable_sort(x, *a, **k):
        dim = a[0] if a else k.get("dim", -1)
        desc = a[1] if len(a) > 1 else k.get("descending", False)
        return _sort(x, dim=dim, descending=desc, stable=True)
    torch.sort = stable_sort
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        det = ref_od.ObjectDetector(CLASSES, mode="rpntrain", use_resnet=False)
    sd = det.state_dict()
    det.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=6))
    det.train()
    for mod_ in det.modules():
        if isinstance(mod_, torch.nn.Dropout):
            mod_.p = 0.0
    nb = make_inputs(seed=14, boxes=10, rels=4)
    _, inds, _, _ = OH.anchor_target_layer(nb["gt_boxes"], (592, 592), rng=np.random.RandomState(2))
    tai = np.column_stack((np.zeros(inds.shape[0], dtype=np.int64), inds)).astype(np.int64)
    np.random.seed(31)
    t = torch.from_numpy
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = det(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), None, None, t(tai))
    torch.sort = _sort
    dout = {"train_anchor_inds": tai}
    for k in ("od_obj_dists", "od_box_deltas", "od_obj_labels", "od_box_targets", "od_box_priors", "rpn_scores", "rpn_box_deltas"):
        dout[k] = getattr(r, k).detach().numpy()
    print("detector train ok:", {k: v.shape for k, v in dout.items()}, "fg rois", int((dout["od_obj_labels"] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "reference_detector_train.npz"), **dout)

    # ---- SGCls TRAINING forward (models/train_rels.py:118-141): relation sampling with the numpy RNG, training-mode
    # BatchNorm, teacher-forced decoder, both cross-entropies. Dropout probabilities are set to 0 on both sides.
    out = {}
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = RelModel(CLASSES, RELS, mode="sgcls", num_gpus=1, require_overlap_det=True, use_resnet=False,
                     use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                     rec_dropout=0.0, **KW)
    sd = m.state_dict()
    m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
    m.train()
    for p_ in m.detector.parameters():                 # models/train_rels.py:51-52
        p_.requires_grad = False
    for mod_ in m.modules():
        if isinstance(mod_, (torch.nn.Dropout, torch.nn.AlphaDropout)):
            mod_.p = 0.0
    nb = make_inputs(seed=12, boxes=14, rels=9)
    np.random.seed(21)
    t = torch.from_numpy
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    import torch.nn.functional as F
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    for k in ("rm_obj_dists", "rm_obj_labels", "rel_dists", "rel_labels"):
        out["train_" + k] = getattr(res, k).detach().numpy()
    out["train_loss"] = np.array(float(loss))
    loss.backward()                                     # the reference's own graph, oracle autograd inside the LSTM stand-in
    names, norms, samples = [], [], []
    for k, p_ in m.named_parameters():
        if p_.grad is not None:
            gflat = p_.grad.reshape(-1)
            idx = (torch.arange(16) * (gflat.numel() - 1)) // 15
            names.append(k); norms.append(float(gflat.double().norm())); samples.append(gflat[idx].numpy())
    out["train_grad_names"], out["train_grad_norms"] = np.array(names), np.array(norms)
    out["train_grad_samples"] = np.stack(samples)
    out["train_bn_running_mean"] = m.union_boxes.conv[2].running_mean.numpy().copy()
    print("sgcls train ok: loss %.5f, rel_labels %s (%d fg)" % (float(loss), tuple(res.rel_labels.shape),
                                                                 int((res.rel_labels[:, -1] > 0).sum())))
    # ---- SGDet TRAINING forward (scripts/refine_for_detection.sh): detections from the frozen detector, IoU relabelling,
    # rel_assignments on the detected boxes (numpy RNG), decoder teacher-forced with background labels, both losses
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = RelModel(CLASSES, RELS, mode="sgdet", num_gpus=1, require_overlap_det=True, use_resnet=False,
                     use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                     rec_dropout=0.0, thresh=0.0, **KW)
    sd = m.state_dict()
    m.load_state_dict(synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3))
    m.train()
    for p_ in m.detector.parameters():
        p_.requires_grad = False
    for mod_ in m.modules():
        if isinstance(mod_, (torch.nn.Dropout, torch.nn.AlphaDropout)):
            mod_.p = 0.0
    nb = make_inputs(seed=11)
    from oracle import host as OH2
    _, inds, _, _ = OH2.anchor_target_layer(nb["gt_boxes"], (592, 592), rng=np.random.RandomState(2))
    tai = np.column_stack((np.zeros(inds.shape[0], dtype=np.int64), inds)).astype(np.int64)
    # With random weights no detection overlaps the random GT boxes. The detections do not depend on the GT (RPN path), so a
    # first pass collects them and the GT of the fixture is built FROM them: 10 detections (jittered) become the GT boxes.
    seen = {}
    hk = m.detector.register_forward_hook(lambda mod, inp, outp: seen.__setitem__("res", outp))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(40)
        m(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]), None, t(tai))
    hk.remove()
    priors = seen["res"].rm_box_priors.detach().numpy()
    rng = np.random.RandomState(5)
    pick = np.arange(0, priors.shape[0], max(1, priors.shape[0] // 10))[:10]
    gt_boxes2 = np.clip(priors[pick] + rng.uniform(-2, 2, (len(pick), 4)), 0, 591).astype(np.float32)
    gt_classes2 = np.stack([np.zeros(len(pick)), rng.randint(1, 151, len(pick))], 1).astype(np.int64)
    pairs = [(a, b) for a in range(len(pick)) for b in range(len(pick)) if a != b]
    sel = np.sort(rng.choice(len(pairs), 12, replace=False))
    gt_rels2 = np.array([[0, pairs[k][0], pairs[k][1], rng.randint(1, 51)] for k in sel], dtype=np.int64)
    out["sgdet_train_gt_boxes"], out["sgdet_train_gt_classes"], out["sgdet_train_gt_rels"] = gt_boxes2, gt_classes2, gt_rels2
    np.random.seed(41)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m(t(nb["imgs"]), nb["im_sizes"], 0, t(gt_boxes2), t(gt_classes2), t(gt_rels2), None, t(tai))
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    for k in ("rm_obj_dists", "rm_obj_labels", "rel_dists", "rel_labels"):
        out["sgdet_train_" + k] = getattr(res, k).detach().numpy()
    out["sgdet_train_loss"] = np.array(float(loss.detach()))
    out["sgdet_train_anchor_inds"] = tai
    print("sgdet train ok: loss %.5f, %d detections (%d labelled), rel_labels %s (%d fg)" % (
        float(loss.detach()), res.rm_obj_labels.size(0), int((res.rm_obj_labels > 0).sum()), tuple(res.rel_labels.shape),
        int((res.rel_labels[:, -1] > 0).sum())))
    np.savez_compressed(os.path.join(HERE, "reference_model_train.npz"), **out)
    return RelModel


if __name__ == "__main__":
    main()
