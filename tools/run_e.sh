mkdir -p gpurun_out
TAG=r2e
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print(d['value'], d['ms_per_step'], d['roofline']['per_kernel'], d['e2e']['value'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
# one full ncu capture of the three wide kernels at a reduced batch (the capture of round-2a hung: shared-barrier race, fixed)
CMD2="python tools/run_wide_once.py cfg3 4736"
timeout 120 $CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc5_wide -s 3 -c 3 -o gpurun_out/prof_wide_$TAG $CMD2 > gpurun_out/ncu2_$TAG.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu2_$TAG.log
timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $? :: $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
