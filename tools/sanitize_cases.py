#!/usr/bin/env python
"""One small forward + backward per kernel family, as the target of compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py wide
families: wide (tcgen05 wide kernels, positions), wide_dense (same kernels, dense 0/1 GSO), n8 (tcgen05 kernels of the 8-node
shape), tile (mma.sync tile kernels, weighted dense GSO), csr (one-launch CSR build + fused CSR forward / backward), gso (dense
GSO builder)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gnnfc
fam = sys.argv[1] if len(sys.argv) > 1 else "wide"
dev = "cuda"
torch.manual_seed(0)
def run(B, N, G, F, K, src):
    m = gnnfc.GraphFilterBatch(G, F, K, activation="leaky_relu").to(dev)
    pos = torch.rand(B, N, 2, device=dev) * (N ** 0.5) * 1.3
    if src == "pos": m.addPositions(pos, 2.0, "binary_le")
    elif src == "norm": m.addPositions(pos, 2.0, "sym_norm_lt")
    elif src == "dense01": m.addGSO((torch.rand(B, 1, N, N, device=dev) < 0.3).float())
    elif src == "densew": m.addGSO(torch.rand(B, 1, N, N, device=dev) * (torch.rand(B, 1, N, N, device=dev) < 0.3))
    elif src == "csr": m.addSparseGSO(pos, 2.0, "binary_le", max_degree=64)
    x = torch.randn(B, G, N, device=dev, requires_grad=True)
    y = m(x)
    y.square().mean().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all() and torch.isfinite(m.weight.grad).all()
    print(fam, src, (B, N, G, F, K), "path", gnnfc._cabi.last_path(), "ok", flush=True)
if fam == "wide":
    run(300, 64, 128, 128, 3, "pos"); run(35, 12, 128, 64, 2, "norm")
elif fam == "wide_dense":
    run(40, 12, 128, 128, 3, "dense01")
elif fam == "n8":
    run(200, 8, 32, 32, 3, "pos"); run(200, 8, 32, 32, 3, "norm")
elif fam == "tile":
    run(50, 12, 128, 128, 3, "densew"); run(37, 20, 64, 48, 2, "pos")
elif fam == "csr":
    run(3, 300, 32, 32, 4, "csr"); m = gnnfc.build_csr(torch.rand(2, 1024, 2, device=dev) * 28.3, 2.0, "sym_norm_lt"); torch.cuda.synchronize(); print("csr build ok")
elif fam == "gso":
    a, S = gnnfc.build_gso(torch.rand(100, 12, 2, device=dev) * 6, 2.0, "sym_norm_lt"); torch.cuda.synchronize(); print("gso ok")
