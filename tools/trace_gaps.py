"""Where is the GPU idle inside one SGCls training step? torch.profiler (CUPTI) trace of steady steps; for the last step:
union of kernel intervals per stream, idle gaps on the compute stream sorted by length with the kernels before / after.
    python tools/trace_gaps.py > gpurun_out/r02_trace_gaps.log"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dataloaders.synthetic import make_numpy_batch, SyntheticBlob
# under torchrun (one rank per GPU) every rank steps, rank 0 alone profiles and prints: the data-parallel timeline
from lib.data_parallel import init_from_env
rank, world, local = init_from_env()
if rank != 0:
    sys.stdout = open(os.devnull, "w")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
model = bench.build_model(dev); opt = bench.get_optim(model, 6e-3)
blobs = [SyntheticBlob(make_numpy_batch(6, seed=i + 10 * rank, image_offset=0), dev) for i in range(3)]
for b in blobs:
    b.scatter()
for i in range(8):
    bench.train_step(model, opt, None, fwd_tuple=blobs[i % 3][0])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        torch.cuda.nvtx.range_push("step")
        bench.train_step(model, opt, None, fwd_tuple=blobs[i % 3][0])
        torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_rank%d.json" % rank)
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0, t1 = ks[0]["ts"], ks[-1]["ts"] + ks[-1]["dur"]
# steps are separated by the float(loss) sync: find the two largest CPU-visible boundaries via the loss D2H memcpy
d2h = [e for e in ks if e.get("cat") == "gpu_memcpy" and "DtoH" in e.get("name", "")]
bounds = [e["ts"] + e["dur"] for e in d2h][-4:]
print("kernels+copies", len(ks), "span ms", (t1 - t0) / 1e3, "DtoH copies", len(d2h))
if len(bounds) >= 2:
    a, b = bounds[-2], bounds[-1]
else:
    a, b = t0 + 2 * (t1 - t0) / 3, t1
step = [e for e in ks if a <= e["ts"] < b]
streams = {}
for e in step:
    streams.setdefault(e["args"].get("stream", 0), []).append(e)
print("last step: %.3f ms, %d gpu events, streams: %s" % ((b - a) / 1e3, len(step), {k: len(v) for k, v in streams.items()}))
main = max(streams.values(), key=len)
busy = sum(e["dur"] for e in main)
gaps = []
prev_end = a
for i, e in enumerate(main):
    if e["ts"] > prev_end:
        gaps.append((e["ts"] - prev_end, main[i - 1]["name"][:60] if i else "<step start>", e["name"][:60], (e["ts"] - a) / 1e3))
    prev_end = max(prev_end, e["ts"] + e["dur"])
print("compute stream: busy %.3f ms, idle %.3f ms in %d gaps" % (busy / 1e3, sum(g[0] for g in gaps) / 1e3, len(gaps)))
other = [v for v in streams.values() if v is not main]
for v in other:
    print("  side stream: %d events, busy %.3f ms, from %.3f to %.3f ms" % (len(v), sum(e["dur"] for e in v) / 1e3, (v[0]["ts"] - a) / 1e3, (v[-1]["ts"] + v[-1]["dur"] - a) / 1e3))
    for e in sorted(v, key=lambda e: -e["dur"])[:12]:
        print("      %8.1f us @ %6.2f ms  %s" % (e["dur"], (e["ts"] - a) / 1e3, e["name"][:70]))
gaps.sort(reverse=True)
print("largest gaps (us, at ms into the step):")
for g in gaps[:25]:
    print("  %7.1f us @ %6.2f ms  after [%s]  before [%s]" % (g[0], g[3], g[1], g[2]))
hist = [0] * 6
for g in gaps:
    hist[min(5, int(g[0] // 10))] += g[0]
print("idle by gap length (us): <10: %.0f, 10-20: %.0f, 20-30: %.0f, 30-40: %.0f, 40-50: %.0f, >=50: %.0f" % tuple(hist))
# idle per 1-ms window of the step
import collections
win = collections.Counter()
for g in gaps:
    win[int(g[3])] += g[0]
print("idle us per 1-ms window:", [int(win[i]) for i in range(int((b - a) / 1e3) + 1)])
