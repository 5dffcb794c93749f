mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "wide or cfg3 or model_level or rollout or stacked or cfg1 or config_dense or dp_two or partial" > gpurun_out/pytest_wide_r2c.log 2>&1
echo "pytest wide exit $? :: $(tail -1 gpurun_out/pytest_wide_r2c.log)"
timeout 300 python tools/time_wide.py cfg3 > gpurun_out/time_wide_r2c.log 2>&1; cat gpurun_out/time_wide_r2c.log
timeout 300 python tools/time_wide.py cfg4 >> gpurun_out/time_wide_r2c.log 2>&1; tail -1 gpurun_out/time_wide_r2c.log
for k in dx dh; do
GFC_LIB=$PWD/gnn-formation-control_b200/libgfc_timeline.so timeout 300 python tools/wide_clocks.py cfg3 2368 $k 0 500 > gpurun_out/timeline_${k}_r2c.log 2>&1
echo "timeline $k exit $?"
done
