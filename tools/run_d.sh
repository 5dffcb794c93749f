mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 100 --timeout-method thread -k "many_tiles" > gpurun_out/pytest_mt.log 2>&1
echo "pytest exit $? :: $(tail -3 gpurun_out/pytest_mt.log | tr '\n' '|')"
