#!/bin/bash
# What the round-end driver runs, in one call, plus the profiling evidence:
#   smoke + pytest -m gpu + bench (cfg3 headline, both arms) [+ ncu launch list of the bench command + full capture of the wide kernels]
#   tools/gpu_bench_profile.sh [ncu] [tag]
mkdir -p gpurun_out
TAG=${2:-run}
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $? :: $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 150 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $? :: $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 1200 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "bench ref exit $?"
if [ "$1" == "ncu" ]; then
CMD="python bench.py --steps 3 --warmup 3 --no-graph --no-extra --cpu-budget 0.5"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "ncu launches exit $?"
CMD2="python tools/run_wide_once.py cfg3 4736"
timeout 120 $CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc5_wide -s 3 -c 3 -o gpurun_out/prof_wide_$TAG $CMD2 > gpurun_out/ncu2_$TAG.log 2>&1
echo "ncu full exit $?"
fi
