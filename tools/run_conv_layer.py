"""Launch one VGG-shaped 3x3 conv a few times (for ncu): python tools/run_conv_layer.py S Cin Cout pair_mode halo_mode [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import motifs_cabi as C
from lib import tc_ops
S, Ci, Co, pm, hm = [int(a) for a in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device("cuda:0")
lib = C.load()
lib.mb200_gemm_set_pair_mode(pm); lib.mb200_conv_set_halo_mode(hm)
conv = torch.nn.Conv2d(Ci, Co, 3, padding=1).to(dev)
xh = torch.randn(6, S, S, Ci, device=dev).bfloat16(); xl = (torch.randn(6, S, S, Ci, device=dev) * 1e-3).bfloat16()
for _ in range(reps):
    tc_ops.conv3x3_relu((xh, xl), 6, S, S, Ci, conv, want_f32=False, want_split=True)
torch.cuda.synchronize()
print("done")
