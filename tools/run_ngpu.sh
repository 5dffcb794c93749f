#!/bin/bash
# bench.py on N GPUs of one box (weak scaling, cfg3): tools/run_ngpu.sh N tag
N=${1:-4}; TAG=${2:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --no-extra --cpu-budget 1 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
echo "bench ${N}gpu exit $?"; grep "^{" gpurun_out/bench_${N}gpu_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['dp_check'], d['e2e']['value'])"
