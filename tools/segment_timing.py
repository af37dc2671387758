"""Per-segment host-enqueue vs device time of one SGCls training step (diagnostic, not a benchmark)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch
import bench
from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
from torch.nn import functional as F

dev = torch.device("cuda:0")
model = bench.build_model(dev); opt = bench.get_optim(model, 6e-3)
blob = SyntheticBlob(make_numpy_batch(6, seed=0), dev); blob.scatter()
for _ in range(3):
    bench.train_step(model, opt, None, fwd_tuple=blob[0])
torch.cuda.synchronize()
seg = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        e = seg.setdefault(label, [0.0, 0.0]); e[0] += (t1 - t0) * 1e3; e[1] += (t2 - t0) * 1e3
        return r
    setattr(obj, name, g)
wrap(model.detector, "forward", "detector.forward (VGG+roi+fc)")
wrap(model, "obj_feature_map", "obj_feature_map")
wrap(model.context, "forward", "context (LSTMs+decoder)")
wrap(model, "visual_rep", "visual_rep (union+fc)")
N = 5
tot = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
for _ in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = model(*blob[0])
    loss = F.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + F.cross_entropy(res.rel_dists, res.rel_labels[:, -1])
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    opt.all_reduce_grads(); opt.step()
    t5 = time.perf_counter(); torch.cuda.synchronize(); t6 = time.perf_counter()
    for i, v in enumerate([t1 - t0, t2 - t0, t3 - t2, t4 - t2, t5 - t4, t6 - t4]):
        tot[i] += v * 1e3
print("segment: host-enqueue ms / total ms (with syncs)")
for k, v in seg.items():
    print("  %-34s %7.2f / %7.2f" % (k, v[0] / N, v[1] / N))
print("  %-34s %7.2f / %7.2f" % ("forward total", tot[0] / N, tot[1] / N))
print("  %-34s %7.2f / %7.2f" % ("backward", tot[2] / N, tot[3] / N))
print("  %-34s %7.2f / %7.2f" % ("allreduce+optimizer", tot[4] / N, tot[5] / N))
