"""Operator microbenchmarks on one B200 (BASELINE.json configs 4 and 5): this library's kernels
beside the reference's kernels compiled unmodified for sm_100a (oracle/_ref). CUDA-event timing,
L2 flushed between iterations. Writes gpurun_out/microbench_ops.json.

    python tools/microbench_ops.py
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import motifs_cabi as C  # noqa: E402
from oracle import ref_loader  # noqa: E402  (bench-only: the reference kernels, as the "before")

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # 256 MB > 126 MB L2


def timeit(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def rand_rois(rng, n, batch):
    x1 = rng.uniform(0, 400, n); y1 = rng.uniform(0, 400, n)
    w = rng.uniform(32, 190, n); h = rng.uniform(32, 190, n)
    return np.concatenate([rng.randint(0, batch, (n, 1)), np.stack([x1, y1, np.minimum(x1 + w, 591), np.minimum(y1 + h, 591)], 1)], 1).astype(np.float32)


def bench_roi_align(res):
    from lib.fpn.roi_align.functions.roi_align import normalize_rois
    ref = ref_loader.ref_kernels()
    pk, kind = peaks()
    rng = np.random.RandomState(0)
    rows = []
    for B, Cn, N in [(1, 512, 1024), (6, 512, 1024), (1, 512, 128), (1, 512, 8192), (1, 256, 1024), (1, 1024, 1024), (6, 512, 1536)]:
        feat = torch.randn(B, Cn, 37, 37, device=dev)
        rois = torch.from_numpy(rand_rois(rng, N, B)).to(dev)
        rn = normalize_rois(rois, 37, 37, 1 / 16)
        out = torch.empty(N, Cn, 7, 7, device=dev)
        st = C.cur_stream()
        lib = C.load()
        mine = timeit(lambda: lib.ROIAlignForwardLaucher(C.ptr(feat), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, 0.0, C.ptr(out), st))
        # algorithmic bytes: output write + each feature element once (at most) + rois
        alg = N * Cn * 49 * 4 + min(B * Cn * 37 * 37 * 4, N * Cn * 37 * 37 * 4) + N * 20
        row = {"B": B, "C": Cn, "N": N, "alg_bytes": alg, "us_median": mine[0], "us_min": mine[1],
               "GBs": alg / mine[0] / 1e3, "frac_of_%s_hbm" % kind: alg / mine[0] / 1e3 / pk["hbm_gbs"]}
        if ref is not None:
            theirs = timeit(lambda: ref.ROIAlignForwardLaucher(C.ptr(feat), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, 0.0, C.ptr(out), st))
            row["ref_kernel_us_median"] = theirs[0]
            row["ref_kernel_GBs"] = alg / theirs[0] / 1e3
        fn = feat.permute(0, 2, 3, 1).contiguous()
        out2 = torch.empty(N, 49, Cn, device=dev)
        nh = timeit(lambda: lib.mb200_roi_align_forward_nhwc(C.ptr(fn), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, 0.0, C.ptr(out2), st))
        row["nhwc_us_median"] = nh[0]
        row["nhwc_GBs"] = alg / nh[0] / 1e3
        row["nhwc_frac_of_%s_hbm" % kind] = alg / nh[0] / 1e3 / pk["hbm_gbs"]
        # the layout the pipeline consumes: NHWC feature map in, [N, C, 7, 7] out (fc6's K order)
        nc = timeit(lambda: lib.mb200_roi_align_forward_nhwc_to_nchw(C.ptr(fn), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, 0.0, C.ptr(out), st))
        row["nhwc_to_nchw_us_median"] = nc[0]
        row["nhwc_to_nchw_frac_of_%s_hbm" % kind] = alg / nc[0] / 1e3 / pk["hbm_gbs"]
        g = torch.randn(N, Cn, 7, 7, device=dev); gi = torch.zeros(B, Cn, 37, 37, device=dev)
        bw = timeit(lambda: lib.ROIAlignBackwardLaucher(C.ptr(g), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, C.ptr(gi), st))
        row["bwd_us_median"] = bw[0]
        if ref is not None:
            row["ref_bwd_us_median"] = timeit(lambda: ref.ROIAlignBackwardLaucher(C.ptr(g), C.ptr(rn), N, B, 37, 37, 7, 7, Cn, C.ptr(gi), st))[0]
        rows.append(row)
        print("roi_align", row, flush=True)
    res["roi_align"] = rows


def bench_nms(res):
    ref = ref_loader.ref_kernels()
    rng = np.random.RandomState(1)
    rows = []
    for n, thr in [(6000, 0.7), (1000, 0.3), (12000, 0.7)]:
        x1 = rng.uniform(0, 400, n); y1 = rng.uniform(0, 400, n)
        b = np.stack([x1, y1, x1 + rng.uniform(16, 190, n), y1 + rng.uniform(16, 190, n)], 1).astype(np.float32)
        bs = torch.from_numpy(b).to(dev)
        keep = (ctypes.c_int * n)()
        lib = C.load()
        import time
        def host_timed(fn, iters=10):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(iters):
                t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
            return float(np.median(ts))
        row = {"n": n, "thresh": thr, "drop_in_us": host_timed(lambda: lib.ApplyNMSGPU(keep, C.ptr(bs), n, thr, 0))}
        from lib.fpn.nms.functions.nms import nms_segments
        row["device_only_us"] = timeit(lambda: nms_segments(bs, [n], thr), flush=False)[0]
        if ref is not None:
            row["ref_us"] = host_timed(lambda: ref.ApplyNMSGPU(keep, C.ptr(bs), n, thr, 0))
        rows.append(row)
        print("nms", row, flush=True)
    res["nms"] = rows


def bench_lstm(res, quick):
    """BASELINE configs[4]: H=512, L=2, B=256, T in {32,...}; plus the real SGCls shapes (B=6, T=20).
    This library: the AlternatingHighwayLSTM module (tcgen05 hoisted projections + persistent recurrence),
    forward and forward+backward. Reference: its kernels compiled unmodified (cuBLAS per step), forward."""
    from torch.nn.utils.rnn import pack_padded_sequence
    from lib.lstm.highway_lstm_cuda.alternating_highway_lstm import AlternatingHighwayLSTM
    ref = ref_loader.ref_kernels()
    rows = []
    cfgs = [(20, 6, 4424, 512, 2), (20, 6, 712, 512, 4), (32, 256, 712, 512, 2), (64, 256, 4424, 512, 2)]
    if not quick:
        cfgs += [(128, 256, 712, 512, 2), (256, 256, 712, 512, 2)]
    for T, B, In, H, L in cfgs:
        torch.manual_seed(0)
        m = AlternatingHighwayLSTM(In, H, L).to(dev).train()
        x = torch.randn(T, B, In, device=dev, requires_grad=True)
        lengths = [T] * B
        drop = torch.ones(L, B, H, device=dev)

        def fwd():
            with torch.no_grad():
                m(pack_padded_sequence(x.detach(), lengths), dropout_weights=drop)

        def fwd_bwd():
            out, _ = m(pack_padded_sequence(x, lengths), dropout_weights=drop)
            out.data.sum().backward()
            m.zero_grad(set_to_none=True); x.grad = None
        t_f = timeit(fwd, iters=5, warmup=2, flush=False)
        t_fb = timeit(fwd_bwd, iters=5, warmup=2, flush=False)
        flops = sum(T * (2 * B * (In if l == 0 else H) * 6 * H + 2 * B * H * 5 * H) for l in range(L))
        row = {"T": T, "B": B, "In": In, "H": H, "L": L, "fwd_us": t_f[0], "fwd_bwd_us": t_fb[0],
               "fwd_TFLOPs": flops / t_f[0] / 1e6}
        if ref is not None:
            len_host = (ctypes.c_int * B)(*lengths)
            h = torch.zeros(L, T + 1, B, H, device=dev); c = torch.zeros(L, T + 1, B, H, device=dev)
            gates = torch.empty(L, T, B, 6 * H, device=dev)
            cublas = ctypes.CDLL("libcublas.so.12"); handle = ctypes.c_void_p(); cublas.cublasCreate_v2(ctypes.byref(handle))
            ti = torch.zeros(B, 6 * H, device=dev); th = torch.zeros(B, 5 * H, device=dev)
            w, bias = m.weight.detach(), m.bias.detach()
            xd = x.detach()

            def theirs():
                ref.highway_lstm_forward_ongpu(In, H, B, L, T, C.ptr(xd), len_host, C.ptr(h), C.ptr(c), C.ptr(ti), C.ptr(th),
                                               C.ptr(w), C.ptr(bias), C.ptr(drop), C.ptr(gates), 1, C.cur_stream(), handle)
            row["ref_fwd_us"] = timeit(theirs, iters=3, warmup=1, flush=False)[0]
        rows.append(row)
        print("lstm", row, flush=True)
    res["lstm"] = rows


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    res = {"device": torch.cuda.get_device_name(0), "peaks": peaks()[0], "peaks_kind": peaks()[1]}
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    outname = ([a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")] or ["microbench_ops.json"])[0]
    for name, fn in [("roi", bench_roi_align), ("nms", bench_nms), ("lstm", lambda r: bench_lstm(r, quick))]:
        if only and name not in only:
            continue
        try:
            fn(res)
        except Exception as e:  # keep going: partial results are still useful
            import traceback; traceback.print_exc()
            res[name + "_error"] = repr(e)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", outname), "w"), indent=1)
