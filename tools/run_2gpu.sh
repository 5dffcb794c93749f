mkdir -p gpurun_out
TAG=r2j
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 150 --timeout-method thread -k "two_gpus" -rA > gpurun_out/pytest_2gpu_$TAG.log 2>&1
echo "pytest two_gpus exit $? :: $(tail -1 gpurun_out/pytest_2gpu_$TAG.log)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_peer_check.py > gpurun_out/dp_peer_$TAG.log 2>&1; echo "peer check exit $?"; grep -v Warning gpurun_out/dp_peer_$TAG.log | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/dp_overlap.py > gpurun_out/dp_overlap_$TAG.log 2>&1; echo "overlap exit $?"; grep -v Warning gpurun_out/dp_overlap_$TAG.log | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --no-extra --cpu-budget 1 > gpurun_out/bench_2gpu_$TAG.json 2> gpurun_out/bench_2gpu_$TAG.err; echo "bench 2gpu exit $?"; tail -c 600 gpurun_out/bench_2gpu_$TAG.json
