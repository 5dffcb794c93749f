mkdir -p gpurun_out
timeout 200 python tools/wide_flush_accuracy.py 1 2 3 4 6 8 > gpurun_out/flush_acc_r2q.log 2>&1; cat gpurun_out/flush_acc_r2q.log
