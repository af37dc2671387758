"""Debug aid (round 2): isolate the in-model LSTM mismatch of the load_state_dict'ed eval models."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lib.rel_model import RelModel
from golden.synthetic_state import synthetic_state, CLASSES, RELS, KW, make_inputs
from oracle.highway_lstm import highway_lstm_forward
from torch.nn.utils.rnn import PackedSequence, pad_packed_sequence


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


cuda = torch.device("cuda:0")
how = sys.argv[1]
torch.manual_seed(0)
prod = RelModel(CLASSES, RELS, mode="predcls", num_gpus=1, require_overlap_det=True, use_resnet=False, use_proposals=False,
                pass_in_obj_feats_to_decoder=False, pass_in_obj_feats_to_edge=False, rec_dropout=0.1, **KW)
sd = prod.state_dict()
state = synthetic_state([(k, tuple(v.shape), v.dtype) for k, v in sd.items()], seed=3)
if how == "load":
    prod.load_state_dict(state)
elif how == "load_no_running":           # everything but BatchNorm running statistics
    prod.load_state_dict({k: (v if "running" not in k else sd[k]) for k, v in state.items()})
elif how == "copy":                      # same values, written with .data.copy_ instead of load_state_dict
    with torch.no_grad():
        for k, v in prod.state_dict().items():
            v.copy_(state[k])
elif how == "frozen":
    prod.load_state_dict(state)
    for p in prod.detector.parameters():
        p.requires_grad = False
prod = prod.to(cuda).eval()
cap = {}


def hook(name):
    def f(mod, inp, out):
        cap[name] = (inp[0], out[0])
    return f


prod.context.obj_ctx_rnn.register_forward_hook(hook("obj"))
prod.context.edge_ctx_rnn.register_forward_hook(hook("edge"))
nb = make_inputs(seed=11)
t = torch.from_numpy
with torch.no_grad():
    prod(t(nb["imgs"]).to(cuda), nb["im_sizes"], 0, t(nb["gt_boxes"]).to(cuda), t(nb["gt_classes"]).to(cuda), t(nb["gt_rels"]).to(cuda))
    torch.cuda.synchronize()
    bad = [k for k, v in prod.state_dict().items() if how in ("load", "frozen", "copy") and not torch.equal(v.cpu(), state[k])]
    print(how, "parameters that differ from the loaded state after forward:", bad[:8], len(bad))
    for name, mod in (("obj", prod.context.obj_ctx_rnn), ("edge", prod.context.edge_ctx_rnn)):
        pin, pout = cap[name]
        padded, lengths = pad_packed_sequence(pin)
        want = highway_lstm_forward(padded.cpu(), [int(l) for l in lengths], mod.weight.detach().cpu(), mod.bias.detach().cpu(),
                                    torch.ones(mod.num_layers, padded.size(1), mod.hidden_size), mod.hidden_size, mod.num_layers)
        got, _ = pad_packed_sequence(pout)
        again, _ = pad_packed_sequence(mod(pin)[0])
        print(how, name, "T,B,In", tuple(padded.shape), "lengths", [int(l) for l in lengths], "training", mod.training,
              "in-model vs CPU recurrence on the SAME input/weights %.3e" % rel(got, want),
              "| re-run standalone %.3e" % rel(again, want), "| input absmax %.3e" % float(padded.abs().max()),
              "out absmax %.3e" % float(want.abs().max()))
