NP=${NP:-2}
CHECK_COMM=ce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29515 tools/check_dp_sharded.py 2>&1 | grep "PASS\|FAIL\|Error\|error" | tail -25
CHECK_COMM=nvls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29516 tools/check_dp_sharded.py 2>&1 | grep "PASS\|FAIL\|Error\|error" | tail -25
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 20 --warmup 3 > gpurun_out/r02_n${NP}_$tag.json 2> gpurun_out/r02_n${NP}_$tag.err; echo "$tag rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_n${NP}_$tag.json'));print('$tag',d['value'],d['ms_per_step'],d['e2e']['value'],d['config'].get('dp_comm'))" || grep -v "NCCL INFO" gpurun_out/r02_n${NP}_$tag.err | tail -15; }
run ce MOTIFS_DP_COMM=auto
run nccl MOTIFS_DP_COMM=nccl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tools/trace_gaps.py > gpurun_out/r02_trace_gaps_n${NP}_ce.log 2> gpurun_out/r02_trace_gaps_n${NP}_ce.err; echo "trace rc=$?"; head -46 gpurun_out/r02_trace_gaps_n${NP}_ce.log | cut -c1-160
