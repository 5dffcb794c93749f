mkdir -p gpurun_out
L=$PWD/gnn-formation-control_b200
run() { # tag lib B [env]
  env $4 GFC_LIB=$2 GFC_B=$3 timeout 45 python tools/time_wide.py cfg3 > gpurun_out/exp_$1_B$3.log 2>&1
  echo "[$1 B=$3] exit $? :: $(tr '\n' '|' < gpurun_out/exp_$1_B$3.log | tail -c 300)"
}
run cur $L/libgfc.so 2368 A=1
run cur $L/libgfc.so 65536 A=1
run nopf $L/libgfc.so 65536 NOPF=1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method thread -k "wide or cfg3 or model_level or rollout or stacked or cfg1 or config_dense or dp_two or partial or reducer or recurrent" > gpurun_out/pytest_wide_r2d.log 2>&1
echo "pytest wide exit $? :: $(tail -3 gpurun_out/pytest_wide_r2d.log | tr '\n' '|')"
for k in fwd dx dh; do
GFC_LIB=$L/libgfc_timeline.so timeout 60 python tools/wide_clocks.py cfg3 2368 $k 0 600 > gpurun_out/timeline_${k}_r2f.log 2>&1
echo "timeline $k exit $?"
done
GFC_B=16384 timeout 45 python tools/time_wide.py cfg4 > gpurun_out/exp_cfg4.log 2>&1; cat gpurun_out/exp_cfg4.log
