# one-GPU validation pass: RoIAlign separable kernel + tensor-core LSTM recurrence
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "roi or lstm" 2>&1 | tail -5
for chw in 0 1; do for mode in 0 1; do
  echo "run_roi chw=$chw cols=$mode"; MOTIFS_ROI_NHWC_COLS=$mode timeout 120 python tools/run_roi.py 1024 $chw
  MOTIFS_ROI_NHWC_COLS=$mode timeout 120 python tools/run_roi.py 8192 $chw
done; done
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "neural-motifs_b200")
import torch, bench
rows = bench.lstm_microbench(torch.device("cuda:0"))
json.dump(rows, open("gpurun_out/r02_lstm_microbench_prefetch.json", "w"), indent=1)
for r in (rows if isinstance(rows, list) else rows.get("rows", [])): print(r)
PY
