"""RoIAlign config 4 (N=1024, C=512, 7x7, B=1) a few times, for ncu: python tools/run_roi.py [N] [chw]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import motifs_cabi as C
from lib.fpn.roi_align.functions.roi_align import normalize_rois
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chw = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
x1 = rng.uniform(0, 400, N); y1 = rng.uniform(0, 400, N); w = rng.uniform(32, 190, N); h = rng.uniform(32, 190, N)
rois = np.concatenate([np.zeros((N, 1)), np.stack([x1, y1, np.minimum(x1 + w, 591), np.minimum(y1 + h, 591)], 1)], 1)
rn = normalize_rois(torch.from_numpy(rois.astype(np.float32)).to(dev), 37, 37, 1 / 16)
feat = torch.randn(1, 37, 37, 512, device=dev)
out = torch.empty(N, 512, 7, 7, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
lib = C.load()
fn = lib.mb200_roi_align_forward_nhwc_to_nchw if chw else lib.mb200_roi_align_forward_nhwc
ts = []
for i in range(8):
    flush.zero_()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); fn(C.ptr(feat), C.ptr(rn), N, 1, 37, 37, 7, 7, 512, 0.0, C.ptr(out), C.cur_stream()); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print("N", N, "us", sorted(ts)[len(ts) // 2])
