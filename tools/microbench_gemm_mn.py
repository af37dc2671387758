"""Weight-gradient GEMM microbenchmark for the next round: dW = dY^T X on the K-major kernel (two split_transpose copies +
gemm_bf16x3) against the experimental MN-major kernel (split_rows pairs consumed directly, csrc/gemm_mn.cu), at the
shapes of the SGCls step (fc6 / fc7 of the union branch, post_lstm, LSTM input projection). CUDA events, L2 flushed.
Writes gpurun_out/microbench_gemm_mn.json.       MOTIFS_GEMM_MN=1 python tools/microbench_gemm_mn.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-motifs_b200"))

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
    from lib import tc_ops
    rows = []
    for name, (K, M, N) in {"fc6_union": (1536, 4096, 25088), "fc7_union": (1536, 4096, 4096),
                            "post_lstm": (120, 8192, 512), "lstm_wi_obj": (120, 3072, 4424),
                            "rel_compress": (1536, 51, 4096)}.items():
        dy = torch.randn(K, M, device=dev); x = torch.randn(K, N, device=dev)
        ref = dy.double().t() @ x.double()

        def k_major():
            return tc_ops.gemm(tc_ops.split_transposed(dy), tc_ops.split_transposed(x))

        def mn_major():
            return tc_ops.gemm_mn(tc_ops.split_rows(dy), tc_ops.split_rows(x))
        err = float((mn_major().double() - ref).abs().max() / ref.abs().max())
        row = {"shape": name, "K": K, "M": M, "N": N, "k_major_us": timeit(k_major), "mn_major_us": timeit(mn_major),
               "mn_relerr_vs_fp64": err, "gflop": 2.0 * K * M * N / 1e9}
        rows.append(row)
        print(row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows},
              open(os.path.join(ROOT, "gpurun_out", "microbench_gemm_mn.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
