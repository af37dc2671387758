"""First-contact check + A/B timing of the CTA-pair (cta_group::2) tcgen05 kernel against the 1-CTA kernel.
    python tools/check_pair_gemm.py            # numerics vs fp64 for GEMM and conv shapes, then timings of the VGG layers
Each case prints one line as soon as it is done (a hang shows where)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))
import motifs_cabi as C
from lib import tc_ops

dev = torch.device("cuda:0")
lib = C.load()


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def gemm_case(M, N, K, mode):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    ref = (x.double() @ w.double().t() + b.double()).clamp_min(0)
    lib.mb200_gemm_set_pair_mode(mode)
    y, ys = tc_ops.gemm(tc_ops.split_rows(x), tc_ops.split_rows(w), bias=b, relu=True, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    return relerr(y, ref), relerr(ys.hi[:, :N].float() + ys.lo[:, :N].float(), ref)


def conv_case(B, H, W, Cin, Cout, mode):
    torch.manual_seed(B * H + W)
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).to(dev)
    x = torch.randn(B, Cin, H, W, device=dev)
    ref = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double(), padding=1).clamp_min(0).permute(0, 2, 3, 1)
    xs = tc_ops.split_rows(x.permute(0, 2, 3, 1).contiguous().view(-1, Cin))
    lib.mb200_gemm_set_pair_mode(mode)
    y, ysp = tc_ops.conv3x3_relu((xs.hi.view(B, H, W, Cin), xs.lo.view(B, H, W, Cin)), B, H, W, Cin, conv, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    return relerr(y, ref), relerr(ysp[0].float() + ysp[1].float(), ref)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {"numerics": [], "timing": []}
ok = True
for (M, N, K) in [(256, 128, 64), (256, 256, 256), (512, 256, 128), (384, 512, 1024), (1536, 4096, 4096), (1000, 600, 712),
                  (257, 640, 712), (2000, 3072, 512), (1536, 51, 4096)]:
    e = gemm_case(M, N, K, 2)
    print("gemm pair", (M, N, K), "relerr %.2e %.2e" % e, flush=True)
    out["numerics"].append(["gemm", M, N, K, e[0], e[1]]); ok &= max(e) < 3e-5
halo_ok = {}
for halo in (0, 2):
    lib.mb200_conv_set_halo_mode(halo)
    good = True
    for (B, H, W, Ci, Co) in [(1, 8, 32, 64, 128), (1, 24, 16, 64, 64), (2, 37, 37, 512, 512), (1, 74, 74, 256, 512), (3, 9, 70, 128, 256),
                              (2, 20, 50, 64, 64), (1, 16, 8, 64, 64), (1, 5, 3, 128, 192)]:
        e = conv_case(B, H, W, Ci, Co, 2)
        print("conv pair halo=%d" % halo, (B, H, W, Ci, Co), "relerr %.2e %.2e" % e, flush=True)
        out["numerics"].append(["conv", halo, B, H, W, Ci, Co, e[0], e[1]]); good &= max(e) < 3e-5
    halo_ok[halo] = good
print("HALO numerics", halo_ok, flush=True)
ok &= halo_ok[0]
HALO = 2 if halo_ok[2] else 0
lib.mb200_conv_set_halo_mode(HALO)
print("NUMERICS", "OK" if ok else "FAIL", flush=True)
if ok and "--time" in sys.argv:
    Bn = 6
    layers = [(592, 64, 64), (296, 64, 128), (296, 128, 128), (148, 128, 256), (148, 256, 256), (74, 256, 512), (74, 512, 512), (37, 512, 512)]
    for (S, Ci, Co) in layers:
        conv = torch.nn.Conv2d(Ci, Co, 3, padding=1).to(dev)
        xh = torch.randn(Bn, S, S, Ci, device=dev).bfloat16(); xl = (torch.randn(Bn, S, S, Ci, device=dev) * 1e-3).bfloat16()
        row = {"layer": [S, Ci, Co], "gflop": 2.0 * Bn * S * S * Co * 9 * Ci / 1e9}
        for mode in (0, 1, 2):
            lib.mb200_gemm_set_pair_mode(mode if mode < 2 else 1)
            lib.mb200_conv_set_halo_mode(0 if mode < 2 else HALO)       # mode0: 1-CTA taps; mode1: pair taps; mode2: pair + halo
            us = timeit(lambda: tc_ops.conv3x3_relu((xh, xl), Bn, S, S, Ci, conv, want_f32=False, want_split=True))
            row["us_mode%d" % mode] = us; row["tflops_mode%d" % mode] = row["gflop"] / us * 1e3
        lib.mb200_gemm_set_pair_mode(1)
        print(json.dumps(row), flush=True); out["timing"].append(row)
    for (M, N, K) in [(1536, 4096, 25088), (1536, 4096, 4096), (1536, 25088, 4096), (4096, 25088, 1536), (120, 4096, 25088), (75264, 512, 2304), (75264, 256, 128)]:
        x = tc_ops.split_rows(torch.randn(M, K, device=dev)); w = tc_ops.split_rows(torch.randn(N, K, device=dev))
        row = {"gemm": [M, N, K], "gflop": 2.0 * M * N * K / 1e9}
        for mode in (0, 1, 2):
            lib.mb200_gemm_set_pair_mode(mode)
            us = timeit(lambda: tc_ops.gemm(x, w))
            row["us_mode%d" % mode] = us; row["tflops_mode%d" % mode] = row["gflop"] / us * 1e3
        print(json.dumps(row), flush=True); out["timing"].append(row)
    lib.mb200_gemm_set_pair_mode(1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_pair_gemm_check.json"), "w"), indent=1)
