"""Where does the host wait for the GPU inside one SGCls training step? torch.cuda.set_sync_debug_mode("warn") makes every
synchronizing torch call warn; the warnings are collected with the Python line that issued them. Also prints the GPU-idle
estimate: step wall time vs the sum of kernel time is in the ncu launch list, this tool only finds the host waits.
    python tools/find_syncs.py [--early]"""
import os, sys, warnings, traceback, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
if "--early" in sys.argv:
    os.environ["MOTIFS_EARLY_HOST_INDS"] = "1"
import torch
import bench
from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
dev = torch.device("cuda:0")
model = bench.build_model(dev); opt = bench.get_optim(model, 6e-3)
blob = SyntheticBlob(make_numpy_batch(6, seed=0), dev); blob.scatter()
for _ in range(4):
    bench.train_step(model, opt, None, fwd_tuple=blob[0])
torch.cuda.synchronize()
sites = collections.Counter()


def showwarning(message, category, filename, lineno, file=None, line=None):
    if "synchroniz" not in str(message).lower():
        return
    st = [f for f in traceback.extract_stack()[:-1] if "/repo/" in f.filename and "find_syncs" not in f.filename]
    sites[" <- ".join("%s:%d" % (os.path.relpath(f.filename, ROOT), f.lineno) for f in st[-3:][::-1])] += 1


warnings.showwarning = showwarning
warnings.simplefilter("always")
torch.cuda.set_sync_debug_mode("warn")
bench.train_step(model, opt, None, fwd_tuple=blob[0])
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
print("synchronizing calls in one step (%s):" % ("MOTIFS_EARLY_HOST_INDS=1" if "--early" in sys.argv else "default"))
for k, v in sites.most_common():
    print("  %3d  %s" % (v, k))
