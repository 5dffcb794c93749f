mkdir -p gpurun_out
timeout 60 python tools/time_wide.py cfg1 > gpurun_out/time_cfg1.log 2>&1; cat gpurun_out/time_cfg1.log
GFC_B=16384 timeout 60 python tools/time_wide.py cfg4 > gpurun_out/time_cfg4.log 2>&1; cat gpurun_out/time_cfg4.log
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method thread -k "wide or cfg3 or model_level or rollout or stacked or cfg1 or config_dense or dp_two or partial or statistics or dense_binary or recurrent" > gpurun_out/pytest_wide_r2d.log 2>&1
echo "pytest wide exit $? :: $(tail -3 gpurun_out/pytest_wide_r2d.log | tr '\n' '|')"
timeout 100 python - <<'PY' > gpurun_out/prof_cfg1.log 2>&1
import sys; sys.path.insert(0,'.')
import torch, gnnfc, bench
for name in ("cfg1", "cfg4"):
    w=dict(bench.WORKLOADS[name]); hp=bench.HotPath(w, torch.device('cuda',0), 1)
    for _ in range(3): hp.step(0)
    torch.cuda.synchronize()
    import torch.profiler as tp
    with tp.profile(activities=[tp.ProfilerActivity.CUDA]) as prof:
        hp.step(0); torch.cuda.synchronize()
    print(name); print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
PY
grep -E "^cfg|gfc|Memset|Self CUDA time" gpurun_out/prof_cfg1.log | cut -c1-75,150-230
