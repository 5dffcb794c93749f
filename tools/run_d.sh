mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 100 --timeout-method thread -k "csr" > gpurun_out/pytest_csr_r2x.log 2>&1
echo "pytest csr exit $? :: $(tail -3 gpurun_out/pytest_csr_r2x.log | tr '\n' '|')"
timeout 120 python tools/time_csr.py > gpurun_out/time_csr_r2x.log 2>&1; grep -E "build_csr|csr_|forward|fwd" gpurun_out/time_csr_r2x.log | cut -c1-70,150-235
