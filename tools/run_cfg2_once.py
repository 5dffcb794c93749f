#!/usr/bin/env python
"""run a few cfg2-shaped steps (for ncu): python tools/run_cfg2_once.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath
w = dict(WORKLOADS["cfg2"]); w["B"] = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
hp = HotPath(w, torch.device("cuda", 0), 2)
for i in range(3):
    hp.step(i % 2)
torch.cuda.synchronize()
print("ok", hp.launches_per_step)
