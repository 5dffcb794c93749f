#!/bin/bash
# pytest (as the driver runs it) + bench + ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? :: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
echo "bench exit $?"; tail -c 3600 gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
if [ "$1" == "ncu" ]; then
CMD="python bench.py --steps 24 --warmup 3 --no-graph --no-extra --cpu-budget 0.5"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc5_n8_bwd -s 4 -c 3 -o gpurun_out/prof_cfg2 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
fi
