mkdir -p gpurun_out
GFC_B=65536 timeout 45 python tools/time_wide.py cfg3 > gpurun_out/exp_cur.log 2>&1; echo "[time_wide] $(tr '\n' '|' < gpurun_out/exp_cur.log)"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method thread -k "wide or cfg3 or model_level or config_dense or partial or dense_binary" > gpurun_out/pytest_wide_r2d.log 2>&1
echo "pytest wide exit $? :: $(tail -3 gpurun_out/pytest_wide_r2d.log | tr '\n' '|')"
