set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_tc_gpu.py -m gpu -x -q -k "roi or flat_sgd or direct_gradient" 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_bg.json 2> gpurun_out/r02_bench_n1_bg.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_n1_bg.json'));print('N1',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
timeout 300 python tools/trace_gaps.py > gpurun_out/r02_trace_gaps_bg.log 2>/dev/null; head -24 gpurun_out/r02_trace_gaps_bg.log | cut -c1-200
