timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/probe_symm.py > gpurun_out/r02_probe_symm_n2.json 2> gpurun_out/r02_probe_symm_n2.err; echo "probe rc=$?"; cat gpurun_out/r02_probe_symm_n2.json; tail -5 gpurun_out/r02_probe_symm_n2.err
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_n2_$tag.json 2> gpurun_out/r02_n2_$tag.err; echo "$tag rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_n2_$tag.json'));print('$tag',d['value'],d['ms_per_step'],d['e2e']['value'])"; }
run res0 MOTIFS_NCCL_SM_RESERVE=0
run res32 MOTIFS_NCCL_SM_RESERVE=32
run ch16res16 MOTIFS_NCCL_SM_RESERVE=16 NCCL_MAX_NCHANNELS=16
