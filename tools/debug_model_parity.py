"""Stage-by-stage error report: product (cuda:0) vs oracle (CPU) on a small SGCls training batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import numpy as np, torch
from tests.model_utils import build_pair, make_masks, to_dev, relerr
from dataloaders.synthetic import make_numpy_batch, to_tuple

cuda = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "sgcls"
B, boxes = 2, 9
prod, orc = build_pair(mode, seed=1)
nb = make_numpy_batch(B, seed=3, boxes_per_img=boxes, rels_per_img=9)
prod = prod.to(cuda).train(); orc.train()
n_obj = B * boxes; n_rel = B * boxes * (boxes - 1)
det, top, ctx = make_masks(n_obj, n_rel, B, seed=5)
prod.detector.dropout_masks = to_dev(det, cuda); prod.dropout_masks = to_dev(top, cuda); prod.context.dropout_masks = to_dev(ctx, cuda)
orc.detector.masks, orc.masks, orc.context.masks = det, top, ctx
prod.detector.rng = np.random.RandomState(11); orc.detector.rng = np.random.RandomState(11)
caps = {"p": {}, "o": {}}
def hook(store, name):
    def f(mod, inp, out):
        o = out
        while isinstance(o, (tuple, list)) or hasattr(o, "data") and not torch.is_tensor(o):
            o = o[0] if isinstance(o, (tuple, list)) else o.data
        store[name] = o.detach().cpu() if torch.is_tensor(o) else o
    return f
for name in ["union_boxes", "union_boxes.conv", "context.obj_ctx_rnn", "context.edge_ctx_rnn", "context.decoder_rnn", "context.pos_embed.0"]:
    dict(prod.named_modules())[name].register_forward_hook(hook(caps["p"], name))
    dict(orc.named_modules())[name].register_forward_hook(hook(caps["o"], name))
rp = prod(*to_tuple(nb, cuda))
ro = orc(torch.from_numpy(nb["imgs"]), nb["im_sizes"], 0, torch.from_numpy(nb["gt_boxes"]), torch.from_numpy(nb["gt_classes"]), torch.from_numpy(nb["gt_rels"]))
print("fmap", relerr(rp.fmap, ro.fmap))
print("od_obj_dists", relerr(rp.od_obj_dists, ro.od_obj_dists))
print("obj_fmap(rel)", relerr(rp.obj_fmap.detach(), ro.obj_fmap.detach()))
for k in caps["p"]:
    a, b = caps["p"][k], caps["o"][k]
    if torch.is_tensor(a) and torch.is_tensor(b) and a.shape == b.shape:
        print(k, relerr(a, b), tuple(a.shape), float(b.abs().max()))
    else:
        print(k, "shape mismatch", getattr(a, "shape", None), getattr(b, "shape", None))
print("rm_obj_dists", relerr(rp.rm_obj_dists.detach(), ro.rm_obj_dists.detach()))
print("rel_dists", relerr(rp.rel_dists.detach(), ro.rel_dists.detach()), float(ro.rel_dists.abs().max()))
