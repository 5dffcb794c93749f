mkdir -p gpurun_out
TAG=r2g
L=$PWD/gnn-formation-control_b200
GFC_B=65536 timeout 45 python tools/time_wide.py cfg3 > gpurun_out/exp_cur_$TAG.log 2>&1; echo "[time_wide] $(tr '\n' '|' < gpurun_out/exp_cur_$TAG.log)"
timeout 500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $? :: $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['roofline']['per_kernel'].items()}, d['e2e']['value'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('breakdown_ms'))
PY
GFC_LIB=$L/libgfc_timeline.so timeout 60 python tools/wide_clocks.py cfg3 2368 dh 0 600 > gpurun_out/timeline_dh_$TAG.log 2>&1
