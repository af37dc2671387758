"""One forward of AlternatingHighwayLSTM at the BASELINE configs[4] shape (B=256, H=512, 2 layers, In=712, T=32), for ncu:
    ncu --set full --import-source on --clock-control none -k regex:lstm_fwd_tc -c 1 -o gpurun_out/ncu_lstm_tc python tools/run_lstm_tc.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch
from torch.nn.utils.rnn import pack_padded_sequence
from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import AlternatingHighwayLSTM
dev = torch.device("cuda:0")
torch.manual_seed(0)
T, B, In = 32, 256, 712
m = AlternatingHighwayLSTM(In, 512, 2, recurrent_dropout_probability=0.1).to(dev).eval()
x = torch.randn(T, B, In, device=dev)
with torch.no_grad():
    for _ in range(2):
        m(pack_padded_sequence(x, [T] * B))
torch.cuda.synchronize()
print("done")
