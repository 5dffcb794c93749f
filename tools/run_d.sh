mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 100 --timeout-method thread -k "csr" > gpurun_out/pytest_csr_r2k.log 2>&1
echo "pytest csr exit $? :: $(tail -3 gpurun_out/pytest_csr_r2k.log | tr '\n' '|')"
timeout 120 python tools/time_csr.py > gpurun_out/time_csr_r2k.log 2>&1; head -6 gpurun_out/time_csr_r2k.log; grep "csr_.*fused" gpurun_out/time_csr_r2k.log | cut -c1-60,150-230
