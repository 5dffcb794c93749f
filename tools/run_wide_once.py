#!/usr/bin/env python
"""Run the cfg3-shaped forward / backward a few times at a reduced batch (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
w = dict(WORKLOADS[name]); w["B"] = B
dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
for _ in range(3):
    hp.step(0)
torch.cuda.synchronize()
print("ok", name, B)
