mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu_r2n.log 2>&1
echo "pytest exit $? :: $(tail -3 gpurun_out/pytest_gpu_r2n.log | tr '\n' '|')"
