#!/usr/bin/env python
"""Where cfg5 (CSR path) spends its time: build / forward / backward, device-timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc, bench
w = bench.CFG5; dev = torch.device("cuda", 0)
B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
pos = torch.from_numpy(bench.make_positions(B, N, w["box"], w["seed"])).to(dev)
x = torch.randn(B, G, N, device=dev).requires_grad_(True)
dY = torch.randn(B, F, N, device=dev)
m = gnnfc.GraphFilterBatch(G, F, K, activation="leaky_relu").to(dev)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
csr = gnnfc.build_csr(pos, 2.0, "binary_le")
print("build_csr (exact sizing, one host read) %.3f ms" % t(lambda: gnnfc.build_csr(pos, 2.0, "binary_le")))
print("build_csr (max_degree=64, sync-free)    %.3f ms" % t(lambda: gnnfc.build_csr(pos, 2.0, "binary_le", max_degree=64)))
print("build_csr (three-pass pair walk)        %.3f ms" % t(lambda: gnnfc.gso.build_csr_three_pass(pos, 2.0, "binary_le")))
m.addSparseGSO(csr)
print("forward   %.3f ms" % t(lambda: m(x)))
y = m(x)
def fb():
    m.zero_grad(set_to_none=True); x.grad = None
    yy = m(x); yy.backward(dY)
print("fwd+bwd   %.3f ms" % t(fb))
import torch.profiler as tp
with tp.profile(activities=[tp.ProfilerActivity.CUDA]) as prof:
    gnnfc.build_csr(pos, 2.0, "binary_le", max_degree=64); gnnfc.build_csr(pos, 2.0, "sym_norm_lt", max_degree=64)
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
