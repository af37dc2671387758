"""Microbenchmark of the union-box mask branch at the SGCls training size
(R = 1536 relation candidates): this library's kernels (MOTIFS_MASKCONV=own) beside the torch/cuDNN
modules with TF32 off in forward (the previous path), forward and forward+backward; plus the fused SGD
kernel's achieved HBM bandwidth. CUDA-event timing, L2 flushed between iterations.
Writes gpurun_out/microbench_maskconv.json.

    python tools/microbench_maskconv.py
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import motifs_cabi as C  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def main():
    from lib import get_union_boxes as gub
    from lib import mask_conv
    res = {"device": torch.cuda.get_device_name(0)}
    R = 1536
    torch.manual_seed(0)
    mod = gub.UnionBoxesAndFeats(pooling_size=7, stride=16, dim=512).to(dev).train()
    net_own, net_cudnn = mod.conv, copy.deepcopy(mod.conv)
    masks = (torch.rand(R, 2, 27, 27, device=dev) - 0.5) * (torch.rand(R, 2, 27, 27, device=dev) > 0.4).float()
    addend = torch.randn(R, 512, 7, 7, device=dev)
    g = torch.randn(R, 512, 7, 7, device=dev)

    def own_fwd():
        with torch.no_grad():
            return mask_conv.mask_conv_net(net_own, masks, addend=addend)

    def own_fb():
        out = mask_conv.mask_conv_net(net_own, masks, addend=addend)
        out.backward(g)
        for p in net_own.parameters():
            p.grad = None

    def cudnn_fwd():
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=True):
            return net_cudnn(masks) + addend

    def cudnn_fb():
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=True):
            out = net_cudnn(masks) + addend
        out.backward(g)
        for p in net_cudnn.parameters():
            p.grad = None

    rows = {}
    for name, fn in [("own_fwd_us", own_fwd), ("own_fwd_bwd_us", own_fb), ("cudnn_fwd_us", cudnn_fwd),
                     ("cudnn_fwd_bwd_us", cudnn_fb)]:
        rows[name] = timeit(fn)
        print(name, rows[name], flush=True)
    rows["R"] = R
    rows["algorithmic_gflop_fwd"] = 2.0 * R * (196 * 256 * 98 + 49 * 512 * 2304) / 1e9
    res["mask_branch"] = rows

    # per-kernel view of one own forward+backward (torch profiler, CUDA time by kernel name)
    try:
        from torch.profiler import profile, ProfilerActivity
        own_fb(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            own_fb(); torch.cuda.synchronize()
        ks = sorted(((e.key, e.self_device_time_total, e.count) for e in prof.key_averages()), key=lambda t: -t[1])[:25]
        res["own_fwd_bwd_kernels_us"] = [{"kernel": k[:80], "us": t, "n": n} for k, t, n in ks]
        for k, t, n in ks:
            print("%9.1f us x%-3d %s" % (t, n, k[:90]), flush=True)
    except Exception as e:  # the profiler is a convenience here
        res["profiler_error"] = repr(e)

    # fused SGD kernel: 24 bytes per parameter (read p, g, m; write p, m, g := 0)
    n = 100 * 1000 * 1000
    p = torch.randn(n, device=dev); gr = torch.randn(n, device=dev); m = torch.zeros(n, device=dev)
    norm = torch.ones(1, device=dev)

    def sgd():
        C.check(C.load().mb200_sgd_momentum_clip(C.ptr(p), C.ptr(gr), C.ptr(m), n, 1e-3, 0.9, 1e-4, C.ptr(norm), 5.0, 0, 1,
                                                 C.cur_stream()), "sgd")
    us = timeit(sgd)
    res["sgd_momentum_clip"] = {"n": n, "us": us, "gbs": 24.0 * n / (us * 1e-6) / 1e9}
    print("sgd", res["sgd_momentum_clip"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench_maskconv.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
