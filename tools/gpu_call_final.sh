# one-GPU validation pass: RoIAlign timing, full GPU suite, smoke, bench, ncu launch list of one steady step
for chw in 0 1; do echo "run_roi chw=$chw: $(timeout 100 python tools/run_roi.py 1024 $chw 2>&1 | tail -1)"; done
echo "== full suite"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 400 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_n1_final.json'));print('N1',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roi_align']['pipeline_nhwc'],d['clocks'])"
echo "== launch list"; timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_step_final.csv python tools/profile_step.py --steps 3 2>&1 | tail -1
