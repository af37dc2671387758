"""Debug aid (round 2): product vs oracle intermediates for the PredCls eval case of tests/test_reference_model_pin_gpu.py."""
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
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


cuda = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "predcls"
prod = RelModel(CLASSES, RELS, mode=mode, num_gpus=1, require_overlap_det=True, use_resnet=False, use_proposals=False,
                pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False, rec_dropout=0.1, **KW)
orc = OM.RelModel(CLASSES, RELS, mode=mode, **KW)
sd = orc.state_dict()
state = synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3)
variant = sys.argv[2] if len(sys.argv) > 2 else "full"
if variant == "default_rnn":        # keep the constructors' LSTM parameters (block-orthogonal, structured bias)
    for k in list(state):
        if "rnn" in k:
            state[k] = prod.state_dict()[k].clone()
prod.load_state_dict(state); orc.load_state_dict(state)
prod = prod.to(cuda).eval(); orc.eval()
prod.keep_last_result = True
cap_p, cap_o = {}, {}


def hook(store, name):
    def f(mod, inp, out):
        while isinstance(out, (tuple, list)) or hasattr(out, "data") and not torch.is_tensor(out):
            out = out[0]
        store[name] = out.detach().float().cpu()
        store[name + "_in"] = [i for i in inp]
    return f


for name in ("context.obj_ctx_rnn", "context.edge_ctx_rnn", "union_boxes", "context.decoder_rnn"):
    for m, store in ((prod, cap_p), (orc, cap_o)):
        sub = m
        try:
            for part in name.split("."):
                sub = getattr(sub, part)
            sub.register_forward_hook(hook(store, name))
        except AttributeError:
            pass
nb = make_inputs(seed=11)
t = torch.from_numpy
with torch.no_grad():
    rp = prod(t(nb["imgs"]).to(cuda), nb["im_sizes"], 0, t(nb["gt_boxes"]).to(cuda), t(nb["gt_classes"]).to(cuda), t(nb["gt_rels"]).to(cuda))
    ro = orc(t(nb["imgs"]), nb["im_sizes"], 0, t(nb["gt_boxes"]), t(nb["gt_classes"]), t(nb["gt_rels"]))
lp, lo = prod.last_result, orc.last_result
for f in ("rm_obj_dists", "obj_fmap", "rel_dists", "obj_preds"):
    a, b = getattr(lp, f, None), getattr(lo, f, None)
    if a is not None and b is not None:
        print(mode, variant, f, tuple(a.shape), "relerr %.3e" % rel(a.float(), b.float()))
for k in cap_p:
    if k.endswith("_in"):
        continue
    if k in cap_o:
        a, b = cap_p[k], cap_o[k]
        print(mode, variant, "module", k, tuple(a.shape), tuple(b.shape), "relerr %.3e" % (rel(a, b) if a.shape == b.shape else -1))
        ip, io = cap_p[k + "_in"][0], cap_o[k + "_in"][0]
        ip = ip.data if hasattr(ip, "batch_sizes") else ip
        if torch.is_tensor(ip) and torch.is_tensor(io) and ip.shape == io.shape:
            print("    input relerr %.3e" % rel(ip.float(), io.float()))
# standalone LSTM, synthetic-style parameters, eval, B=1
from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import AlternatingHighwayLSTM
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
from oracle.highway_lstm import highway_lstm_forward
for (T, B, In, H, L) in ((20, 1, 4424, 512, 2), (20, 1, 712, 512, 4), (20, 3, 712, 512, 4)):
    for style in ("synthetic", "default"):
        torch.manual_seed(0)
        m = AlternatingHighwayLSTM(In, H, L)
        if style == "synthetic":
            with torch.no_grad():
                m.weight.copy_(torch.randn_like(m.weight) * 0.03); m.bias.copy_(torch.randn_like(m.bias) * 0.05)
        x = torch.randn(T, B, In)
        lengths = [T] * B
        with torch.no_grad():
            want = highway_lstm_forward(x, lengths, m.weight.detach(), m.bias.detach(), torch.ones(L, B, H), H, L)
            mc = m.to(cuda).eval()
            out, _ = mc(pack_padded_sequence(x.to(cuda), lengths))
            got, _ = pad_packed_sequence(out, total_length=T)
        print("lstm", (T, B, In, H, L), style, "relerr %.3e" % rel(got, want))
