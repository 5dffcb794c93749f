mkdir -p gpurun_out
GFC_B=65536 timeout 45 python tools/time_wide.py cfg3 > gpurun_out/exp_cur.log 2>&1; echo "[bulk on ] $(tr '\n' '|' < gpurun_out/exp_cur.log)"
NOPF=1 GFC_B=65536 timeout 45 python tools/time_wide.py cfg3 > gpurun_out/exp_cur2.log 2>&1; echo "[bulk off] $(tr '\n' '|' < gpurun_out/exp_cur2.log)"
GFC_B=16384 timeout 45 python tools/time_wide.py cfg4 > gpurun_out/exp_cfg4.log 2>&1; cat gpurun_out/exp_cfg4.log
NOPF=1 GFC_B=16384 timeout 45 python tools/time_wide.py cfg4 > gpurun_out/exp_cfg4.log 2>&1; cat gpurun_out/exp_cfg4.log
