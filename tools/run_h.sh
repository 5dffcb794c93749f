mkdir -p gpurun_out
TAG=r2h
timeout 500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $? :: $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
for fam in wide wide_dense n8 tile csr gso; do
  timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py $fam > gpurun_out/san_memcheck_$fam.log 2>&1
  echo "memcheck $fam exit $? :: $(grep -E 'ERROR SUMMARY|path' gpurun_out/san_memcheck_$fam.log | tr '\n' '|' | tail -c 300)"
done
for fam in n8 tile csr wide; do
  timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py $fam > gpurun_out/san_racecheck_$fam.log 2>&1
  echo "racecheck $fam exit $? :: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/san_racecheck_$fam.log | tr '\n' '|' | tail -c 300)"
done
