"""Debug aid (round 2): product vs oracle intermediates. usage: debug_variant3.py sgdet | <variant name of tests/test_model_variants_gpu.py>"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lib.rel_model import RelModel
from oracle import model as OM
from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    if a.shape != b.shape:
        return "SHAPE %s vs %s" % (tuple(a.shape), tuple(b.shape))
    return "%.3e" % float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


cuda = torch.device("cuda:0")
what = sys.argv[1]
if what == "sgdet":
    mode, kw, pass_in, thresh, seed_in = "sgdet", dict(KW), {}, 0.0, 11
else:
    import test_model_variants_gpu as TV
    mode, kw, pass_in = TV.VARIANTS[what]
    thresh, seed_in = 0.01, 18
flags = dict(pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False); flags.update(pass_in)
prod = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False, use_proposals=False,
                rec_dropout=0.1, thresh=thresh, **flags, **kw)
orc = OM.RelModel(CLASSES, RELS, mode=mode, thresh=thresh, **kw, **pass_in)
sd = orc.state_dict()
state = synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3)
prod.load_state_dict(state); orc.load_state_dict(state)
prod = prod.to(cuda).eval(); orc.eval()
prod.keep_last_result = True
cap_p, cap_o = {}, {}


def hook(store, name):
    def f(mod, inp, out):
        o = out
        while isinstance(o, (tuple, list)):
            o = o[0]
        store[name] = o.detach().float().cpu()
        i = inp[0]
        while isinstance(i, (tuple, list)):
            i = i[0]
        store[name + "_in"] = i.detach().float().cpu() if torch.is_tensor(i) else None
    return f


for name in ("context.obj_ctx_rnn", "context.edge_ctx_rnn", "union_boxes", "context.decoder_rnn", "context.pos_embed"):
    for m, store in ((prod, cap_p), (orc, cap_o)):
        sub = m
        try:
            for part in name.split("."):
                sub = getattr(sub, part)
            sub.register_forward_hook(hook(store, name))
        except AttributeError:
            pass
_vr = prod.visual_rep
prod.visual_rep = lambda *a, **k: cap_p.setdefault("vr", _vr(*a, **k))
if len(orc.roi_fmap) > 2:
    orc.roi_fmap[2].register_forward_hook(lambda m, i, o: cap_o.update(vr=o.detach()))
else:
    orc.roi_fmap[1].register_forward_hook(lambda m, i, o: cap_o.update(vr=o.detach()))
nb = make_inputs(seed=seed_in) if what == "sgdet" else make_inputs(seed=seed_in, boxes=14, rels=5)
t = torch.from_numpy
with torch.no_grad():
    if mode == "sgdet":
        rp = prod(t(nb["imgs"]).to(cuda), nb["im_sizes"], 0)
        ro = orc(t(nb["imgs"]), nb["im_sizes"], 0)
    else:
        rp = prod(t(nb["imgs"]).to(cuda), nb["im_sizes"], 0, t(nb["gt_boxes"]).to(cuda), t(nb["gt_classes"]).to(cuda), t(nb["gt_rels"]).to(cuda))
        ro = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
lp, lo = prod.last_result, orc.last_result
for f in ("od_obj_dists", "rm_obj_dists", "rm_box_priors", "boxes_all", "obj_fmap", "obj_preds", "im_inds", "rel_dists", "obj_scores"):
    a, b = getattr(lp, f, None), getattr(lo, f, None)
    if a is not None and b is not None:
        print(what, f, tuple(a.shape), "relerr", rel(a.float(), b.float()),
              ("equal %.3f" % float((a.cpu() == b).float().mean())) if a.dtype == torch.int64 and a.shape == b.shape else "")
if "vr" in cap_p and "vr" in cap_o:
    print(what, "vr", rel(cap_p["vr"].float(), cap_o["vr"].float()))
for k in cap_p:
    if k.endswith("_in") or k == "vr":
        continue
    if k in cap_o:
        print(what, "module", k, "out relerr", rel(cap_p[k], cap_o[k]), "| in relerr",
              rel(cap_p[k + "_in"], cap_o[k + "_in"]) if cap_p.get(k + "_in") is not None and cap_o.get(k + "_in") is not None else "n/a")
if mode == "sgdet":
    rois_p = getattr(lp, "od_box_priors", None)
    print("num rois product", None if rois_p is None else rois_p.shape, "oracle", lo.rois.shape)
    if rois_p is not None and rois_p.shape[0] == lo.rois.shape[0]:
        print("rois relerr", rel(rois_p, lo.rois[:, 1:]), "max abs", float((rois_p.cpu() - lo.rois[:, 1:]).abs().max()))
    pb, po = np.asarray(rp[0]), np.asarray(ro[0])
    print("final boxes max abs diff", np.abs(pb - po).max() if pb.shape == po.shape else (pb.shape, po.shape),
          "labels equal", (np.asarray(rp[1]) == np.asarray(ro[1])).mean() if pb.shape == po.shape else "n/a")
