"""The reference's CALLER code executed against the product's API surface (SURVEY.md section 8b "Model API").

`models/train_rels.py` (get_optim :57-72, train_batch :118-152) and `models/eval_rels.py` (val_batch :59-86) are parsed from
/root/reference with `ast`, their function definitions are extracted UNEDITED apart from the documented compat shim
below, and are executed here against
  * the product's `RelModel` class — `detector[b]` (`__getitem__` + `Blob.scatter` protocol), `detector.train()/eval()`,
    `named_parameters()` with the `roi_fmap*` prefixes, the training `Result` fields and the eval 5-tuple,
  * the product's `lib.pytorch_misc.clip_grad_norm`, `lib.evaluation.sg_eval.BasicSceneGraphEvaluator`, `config` constants,
  * `dataloaders.synthetic.SyntheticBlob` standing in for `dataloaders/blob.py:Blob`.
The one thing that is NOT the product here is the arithmetic of `forward`: the product has no CPU path (by design) and
/root/reference does not exist on the GPU box, so the two can never meet in one process. `forward` is therefore delegated
to the oracle restatement operating ON THE PRODUCT'S OWN Parameter objects (the oracle modules are re-pointed at them), so
that the callers' `loss.backward()`, `clip_grad_norm(detector.named_parameters())` and `optimizer.step()` act on the
product model. The same step with the product's kernels is held to the oracle on the GPU (tests/test_model_gpu.py, and
tests/test_callers_gpu.py runs the restated caller sequence end to end there).

Compat shim: `x.data[0]` (PyTorch-0.3 indexing of a 0-dim tensor) -> `float(x)`; the `verbose=`
keyword of ReduceLROnPlateau (removed from torch) is dropped. Nothing else is rewritten."""
import ast
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


class _Shim(ast.NodeTransformer):
    def visit_Subscript(self, node):
        self.generic_visit(node)
        if isinstance(node.value, ast.Attribute) and node.value.attr == "data" and \
                isinstance(node.slice, ast.Constant) and node.slice.value == 0:
            return ast.copy_location(ast.Call(func=ast.Name(id="float", ctx=ast.Load()), args=[node.value.value], keywords=[]), node)
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, ast.Name) and node.func.id == "ReduceLROnPlateau":
            node.keywords = [k for k in node.keywords if k.arg != "verbose"]
        return node


def _function(path, name, namespace):
    """exec the (shimmed) definition of top-level function `name` of the reference file into `namespace`."""
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.fix_missing_locations(_Shim().visit(ast.Module(body=[node], type_ignores=[])))
    exec(compile(mod, path, "exec"), namespace)
    return namespace[name]


_BUILT = {}


def _build(mode):
    """Product RelModel (CPU-constructed) whose forward is the oracle's arithmetic on the product's own parameters.
    Built once per module (constructing the three VGG fc stacks and the 1.7 GB synthetic state dominates the run time);
    every caller gets the state reloaded."""
    if mode in _BUILT:
        prod, orc, state = _BUILT[mode]
        prod.load_state_dict(state)
        for p in prod.parameters():
            p.requires_grad = True
            p.grad = None
        return prod, orc, state
    from lib.rel_model import RelModel
    from lib.object_detector import Result
    from oracle import model as OM
    from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW

    class ProductSurface(RelModel):
        def forward(self, x, im_sizes, image_offset, gt_boxes=None, gt_classes=None, gt_rels=None, proposals=None,
                    train_anchor_inds=None, return_fmap=False):
            orc = self.__dict__["_orc"]
            orc.train(self.training)
            out = orc(x, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels)
            if self.training:
                return Result(rm_obj_dists=out.rm_obj_dists, rm_obj_labels=out.rm_obj_labels, rel_dists=out.rel_dists,
                              rel_labels=out.rel_labels, obj_preds=out.obj_preds)
            return out

    prod = ProductSurface(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False,
                          use_proposals=False, pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False,
                          rec_dropout=0.1, **KW)
    orc = OM.RelModel(CLASSES, RELS, mode=mode, **KW)
    sd = orc.state_dict()
    state = synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3)
    prod.load_state_dict(state)
    # re-point every oracle parameter / buffer at the product's object of the same state-dict name
    for name, p in list(prod.named_parameters()) + list(prod.named_buffers()):
        mod = orc
        parts = name.split(".")
        for part in parts[:-1]:
            mod = getattr(mod, part)
        store = mod._parameters if parts[-1] in mod._parameters else mod._buffers
        assert parts[-1] in store, name
        store[parts[-1]] = p
    prod.__dict__["_orc"] = orc                    # (not registered as a submodule: the state dict stays the product's)
    _BUILT[mode] = (prod, orc, state)
    return prod, orc, state


def _ones_masks(orc, n_obj, n_rel, n_img):
    from model_utils import make_masks
    det, top, ctx = make_masks(n_obj, n_rel, n_img, seed=0)
    ones = lambda d: {k: torch.ones_like(v) for k, v in d.items()}
    orc.detector.masks, orc.masks, orc.context.masks = ones(det), ones(top), ones(ctx)


def test_train_rels_get_optim_and_train_batch_run_against_the_product():
    import pandas as pd
    from torch import optim
    from torch.nn import functional as F
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    from lib.pytorch_misc import clip_grad_norm
    from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
    detector, orc, state = _build("sgcls")
    for n, param in detector.detector.named_parameters():            # train_rels.py:51-52
        param.requires_grad = False
    conf = types.SimpleNamespace(adam=False, l2=1e-4, lr=6e-3, clip=5.0, mode="sgcls", num_gpus=1, print_interval=100)
    ns = dict(detector=detector, conf=conf, optim=optim, ReduceLROnPlateau=ReduceLROnPlateau, F=F, pd=pd,
              clip_grad_norm=clip_grad_norm)
    path = os.path.join(REF, "models", "train_rels.py")
    optimizer, scheduler = _function(path, "get_optim", ns)(conf.lr)
    assert [g["lr"] for g in optimizer.param_groups] == [conf.lr / 10.0, conf.lr]
    assert len(optimizer.param_groups[0]["params"]) == 8              # fc6 / fc7 weights + biases of roi_fmap and roi_fmap_obj
    ns["optimizer"] = optimizer
    train_batch = _function(path, "train_batch", ns)
    B, boxes = 1, 6
    nb = make_numpy_batch(B, seed=5, boxes_per_img=boxes, rels_per_img=5)
    _ones_masks(orc, B * boxes, B * boxes * (boxes - 1), B)
    orc.detector.rng = np.random.RandomState(7)
    before = {n: p.detach().clone() for n, p in detector.named_parameters()}
    detector.train()
    res = train_batch(SyntheticBlob(nb, "cpu"), verbose=False)
    assert set(res.index) == {"class_loss", "rel_loss", "total"} and abs(res["total"] - res["class_loss"] - res["rel_loss"]) < 1e-5
    moved = [n for n, p in detector.named_parameters() if not torch.equal(p, before[n])]
    assert moved and all(not n.startswith("detector.") for n in moved)       # the frozen detector did not move
    assert any(n.startswith("roi_fmap_obj.") for n in moved) and any(n.startswith("context.obj_ctx_rnn") for n in moved)
    scheduler.step(0.1); scheduler.step(0.1)                                 # the recipe's plateau schedule accepts it

    # the same step taken by an independent oracle model with torch's own clip + SGD gives the same loss and update
    from oracle import model as OM
    from golden.synthetic_state import CLASSES, RELS, KW
    ref = OM.RelModel(CLASSES, RELS, mode="sgcls", **KW)
    ref.load_state_dict(state); ref.train()
    for p in ref.detector.parameters():
        p.requires_grad = False
    _ones_masks(ref, B * boxes, B * boxes * (boxes - 1), B)
    ref.detector.rng = np.random.RandomState(7)
    fc = [p for n, p in ref.named_parameters() if n.startswith("roi_fmap") and p.requires_grad]
    non_fc = [p for n, p in ref.named_parameters() if not n.startswith("roi_fmap") and p.requires_grad]
    opt = torch.optim.SGD([{"params": fc, "lr": conf.lr / 10.0}, {"params": non_fc}], lr=conf.lr, momentum=0.9, weight_decay=1e-4)
    t = torch.from_numpy
    out = ref(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
    loss = F.cross_entropy(out.rm_obj_dists, out.rm_obj_labels) + F.cross_entropy(out.rel_dists, out.rel_labels[:, -1])
    opt.zero_grad(); loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.grad is not None], 5.0)
    opt.step()
    assert abs(float(loss.detach()) - res["total"]) < 1e-5 * max(1.0, abs(res["total"]))
    want = dict(ref.named_parameters())
    for n, p in detector.named_parameters():
        assert torch.allclose(p, want[n], rtol=1e-5, atol=1e-7), n


def test_eval_rels_val_batch_runs_against_the_product():
    from config import BOX_SCALE, IM_SCALE
    from lib.evaluation.sg_eval import BasicSceneGraphEvaluator
    from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
    detector, orc, _ = _build("sgcls")
    detector.eval()
    nb = make_numpy_batch(1, seed=9, boxes_per_img=7, rels_per_img=6)
    val = types.SimpleNamespace(gt_classes=[nb["gt_classes"][:, 1].copy()],
                                relationships=[nb["gt_rels"][:, 1:].copy()],
                                gt_boxes=[nb["gt_boxes"] * BOX_SCALE / IM_SCALE])
    conf = types.SimpleNamespace(mode="sgcls", num_gpus=1)
    all_pred_entries = []
    ns = dict(detector=detector, conf=conf, val=val, np=np, BOX_SCALE=BOX_SCALE, IM_SCALE=IM_SCALE,
              all_pred_entries=all_pred_entries)
    val_batch = _function(os.path.join(REF, "models", "eval_rels.py"), "val_batch", ns)
    evaluator = BasicSceneGraphEvaluator.all_modes()
    with torch.no_grad():
        val_batch(0, SyntheticBlob(nb, "cpu"), evaluator)
    assert len(all_pred_entries) == 1
    e = all_pred_entries[0]
    assert e["pred_boxes"].shape == (7, 4) and e["pred_rel_inds"].shape == (42, 2) and e["rel_scores"].shape == (42, 51)
    rec = evaluator["sgcls"].result_dict["sgcls_recall"]
    assert all(len(rec[k]) == 1 and 0.0 <= rec[k][0] <= 1.0 for k in (20, 50, 100))
