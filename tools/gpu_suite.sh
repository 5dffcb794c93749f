#!/bin/bash
# Runs the GPU checks group by group (separate processes, so one sticky CUDA
# error does not hide the other groups) and leaves logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
for grp in test_gso test_filter_matches test_output_is test_config_dense test_cfg1_real test_asymmetric test_edge_shapes \
           "test_isolated or test_nin or test_float64 or test_tf32" "test_csr or test_large_dense" \
           test_cfg2_full test_cfg3_full "test_dp_two or test_cabi"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "$grp" \
      > "gpurun_out/pytest_${name}.log" 2>&1
  echo "[$grp] exit $? :: $(tail -1 gpurun_out/pytest_${name}.log)" | tee -a gpurun_out/summary.txt
done
