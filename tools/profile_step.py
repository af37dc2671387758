"""A few SGCls training steps of the bench workload, for ncu (no timing claims made here).
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3
Only the LAST step sits between cudaProfilerStart/Stop: with --profile-from-start off the start-up and the
first steps (weight splits, autotuning) run at native speed and the launch list is exactly one steady step.
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch
import bench
from dataloaders.synthetic import make_numpy_batch, SyntheticBlob

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=2); a = ap.parse_args()
dev = torch.device("cuda:0")
model = bench.build_model(dev)
opt = bench.get_optim(model, 6e-3)
red = None
blob = SyntheticBlob(make_numpy_batch(6, seed=0), dev); blob.scatter()
for i in range(a.steps):
    last = i == a.steps - 1
    if last:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    torch.cuda.nvtx.range_push("step%d" % i)
    bench.train_step(model, opt, red, fwd_tuple=blob[0])
    torch.cuda.nvtx.range_pop()
    if last:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("done")
