#!/usr/bin/env python
"""a few launches of the one-launch CSR builder at the cfg5 shape, as the target of ncu -k regex:csr_build_fused"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc, bench
w = bench.CFG5
pos = torch.from_numpy(bench.make_positions(w["B"], w["N"], w["box"], w["seed"])).cuda()
for _ in range(3):
    csr = gnnfc.build_csr(pos, 2.0, "binary_le", max_degree=64)
torch.cuda.synchronize(); print("ok", csr.check())
