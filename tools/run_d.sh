mkdir -p gpurun_out
L=$PWD/gnn-formation-control_b200
GFC_LIB=$L/libgfc_timeline.so timeout 60 python tools/wide_clocks.py cfg4 16384 fwd 0 500 > gpurun_out/timeline_cfg4_fwd.log 2>&1
echo "timeline exit $?"
