"""cProfile of the host side of a few SGCls training steps (where does Python time go?)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch, bench
from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
dev = torch.device("cuda:0")
model = bench.build_model(dev); opt = bench.get_optim(model, 6e-3)
blob = SyntheticBlob(make_numpy_batch(6, seed=0), dev); blob.scatter()
for _ in range(4):
    bench.train_step(model, opt, None, fwd_tuple=blob[0])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    bench.train_step(model, opt, None, fwd_tuple=blob[0])
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:5000])
