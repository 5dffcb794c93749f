#!/usr/bin/env python
"""A few steps of one bench workload through the C ABI (bench.HotPath, activation mask hand-over included), as the
target of `ncu --metrics gpu__time_duration.sum -k regex:tc5_wide ...`:   tools/run_step_once.py [cfg] [steps] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 3:
    w["B"] = int(sys.argv[3])
hp = bench.HotPath(w, torch.device("cuda", 0), 1)
for _ in range(steps):
    hp.step(0)
torch.cuda.synchronize()
print("ok", hp.launches_per_step, "launches per step")
