"""The "before" for the tensor-core path (SURVEY.md section 2b / BASELINE.md 4.3): torch 2.11's cuDNN / cuBLAS on the 13
VGG16 conv shapes (B = 6, 592x592 input) and on fc6 / fc7 (1536 union boxes and 120 objects), in strict fp32 (the
reference's arithmetic) and with TF32 allowed, beside this library's bf16x3 kernels on the same shapes.
CUDA events, median of 10 after 3 warm-ups, L2 flushed between iterations. Library code is measured here only as a
baseline; nothing on the product path calls it.
    python tools/library_baseline.py        -> gpurun_out/r02_library_baseline.json"""
import json, os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
from lib import tc_ops

dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def set_tf32(on):
    torch.backends.cudnn.allow_tf32 = on
    torch.backends.cuda.matmul.allow_tf32 = on


torch.backends.cudnn.benchmark = True
B = 6
layers = [(592, 3, 64), (592, 64, 64), (296, 64, 128), (296, 128, 128), (148, 128, 256), (148, 256, 256), (148, 256, 256),
          (74, 256, 512), (74, 512, 512), (74, 512, 512), (37, 512, 512), (37, 512, 512), (37, 512, 512)]
rows = []
tot = {"cudnn_fp32": 0.0, "cudnn_tf32": 0.0, "cudnn_tf32_nhwc": 0.0, "own_bf16x3": 0.0}
for (S, Ci, Co) in layers:
    conv = torch.nn.Conv2d(Ci, Co, 3, padding=1).to(dev)
    x = torch.randn(B, Ci, S, S, device=dev)
    gflop = 2.0 * B * S * S * Co * 9 * Ci / 1e9
    row = {"layer": [S, Ci, Co], "gflop": gflop}
    with torch.no_grad():
        set_tf32(False); row["cudnn_fp32_us"] = timeit(lambda: F.conv2d(x, conv.weight, conv.bias, padding=1))
        set_tf32(True); row["cudnn_tf32_us"] = timeit(lambda: F.conv2d(x, conv.weight, conv.bias, padding=1))
        xc = x.contiguous(memory_format=torch.channels_last); wc = conv.weight.detach().contiguous(memory_format=torch.channels_last)
        row["cudnn_tf32_nhwc_us"] = timeit(lambda: F.conv2d(xc, wc, conv.bias, padding=1))
        set_tf32(False)
        if Ci % 64 == 0:
            xs = tc_ops.split_rows(x.permute(0, 2, 3, 1).reshape(-1, Ci).contiguous())
            pair = (xs.hi.view(B, S, S, Ci), xs.lo.view(B, S, S, Ci))
            row["own_bf16x3_us"] = timeit(lambda: tc_ops.conv3x3_relu(pair, B, S, S, Ci, conv, want_f32=False, want_split=True))
        else:
            import motifs_cabi as C
            xh = torch.empty(B, S, S, Co, dtype=torch.bfloat16, device=dev); xl = torch.empty_like(xh)
            w0 = conv.weight.detach().contiguous(); b0 = conv.bias.detach()
            row["own_bf16x3_us"] = timeit(lambda: C.load().mb200_conv3x3_stem_split(C.ptr(x), C.ptr(w0), C.ptr(b0), B, S, S, Co, 1,
                                                                                     C.ptr(xh), C.ptr(xl), C.cur_stream()))
    for k in tot:
        tot[k] += row[k + "_us"]
        row[k + "_tflops"] = gflop / row[k + "_us"] * 1e3
    print(json.dumps(row), flush=True)
    rows.append(row)
    del conv, x
fc = []
for (M, K, N, name) in [(1536, 25088, 4096, "fc6 union"), (1536, 4096, 4096, "fc7 union"), (120, 25088, 4096, "fc6 objects"),
                        (120, 4096, 4096, "fc7 objects")]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    row = {"gemm": name, "MKN": [M, K, N], "gflop": 2.0 * M * N * K / 1e9}
    set_tf32(False); row["cublas_fp32_us"] = timeit(lambda: F.linear(x, w, b))
    set_tf32(True); row["cublas_tf32_us"] = timeit(lambda: F.linear(x, w, b))
    set_tf32(False)
    xs, ws = tc_ops.split_rows(x), tc_ops.split_rows(w)
    row["own_bf16x3_us"] = timeit(lambda: tc_ops.gemm(xs, ws, bias=b))
    row["own_bf16x3_incl_split_us"] = timeit(lambda: tc_ops.gemm(tc_ops.split_rows(x), ws, bias=b))
    for k in ("cublas_fp32", "cublas_tf32", "own_bf16x3"):
        row[k + "_tflops"] = row["gflop"] / row[k + "_us"] * 1e3
    print(json.dumps(row), flush=True)
    fc.append(row)
    del x, w
out = {"what": "torch 2.11 cuDNN/cuBLAS (fp32 strict, TF32 allowed) vs this library's bf16x3 tcgen05 kernels; B=6 VGG16 @592x592",
       "conv": rows, "conv_total_us": tot, "fc": fc}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_library_baseline.json"), "w"), indent=1)
print("TOTAL", json.dumps(tot))
