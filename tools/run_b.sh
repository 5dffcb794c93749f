mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "csr or recurrent or rollout or reducer" > gpurun_out/pytest_new_r2b.log 2>&1
echo "pytest new exit $? :: $(tail -1 gpurun_out/pytest_new_r2b.log)"
timeout 300 python tools/time_csr.py > gpurun_out/time_csr_r2b.log 2>&1; tail -12 gpurun_out/time_csr_r2b.log
for k in fwd dx dh; do
GFC_LIB=$PWD/gnn-formation-control_b200/libgfc_timeline.so timeout 300 python tools/wide_clocks.py cfg3 2368 $k 0 400 > gpurun_out/timeline_${k}_r2b.log 2>&1
echo "timeline $k exit $?"
done
timeout 300 python tools/time_wide.py cfg3 > gpurun_out/time_wide_r2b.log 2>&1; cat gpurun_out/time_wide_r2b.log
