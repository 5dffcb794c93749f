#!/usr/bin/env python
"""one cfg5 (CSR path) forward + backward, for ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc, bench
w = bench.CFG5; dev = torch.device("cuda", 0)
B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
pos = torch.from_numpy(bench.make_positions(B, N, w["box"], w["seed"])).to(dev)
x = torch.randn(B, G, N, device=dev).requires_grad_(True)
dY = torch.randn(B, F, N, device=dev)
m = gnnfc.GraphFilterBatch(G, F, K, activation="leaky_relu").to(dev)
m.addSparseGSO(gnnfc.build_csr(pos, 2.0, "binary_le"))
for _ in range(2):
    m.zero_grad(set_to_none=True); x.grad = None
    m(x).backward(dY)
torch.cuda.synchronize()
print("ok")
