#!/usr/bin/env python
"""Per-launch breakdown of one device-timed step of a bench workload (torch profiler, CUDA activities):
   tools/time_step.py [cfg1|cfg2|cfg3|cfg4] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2: w["B"] = int(sys.argv[2])
dev = torch.device("cuda", 0)
hp = bench.HotPath(w, dev, 2)
for i in range(4): hp.step(i % 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): hp.step(i % 2)
e1.record(); torch.cuda.synchronize()
print("%s B=%d: %.4f ms/step (python loop), %d launches/step" % (name, w["B"], e0.elapsed_time(e1) / 20, hp.launches_per_step))
import torch.profiler as tp
with tp.profile(activities=[tp.ProfilerActivity.CUDA]) as prof:
    for i in range(10): hp.step(i % 2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
