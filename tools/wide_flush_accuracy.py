#!/usr/bin/env python
"""cfg3 at full batch: error of the tcgen05 dH kernel against an fp64 torch reference and its duration, as a
function of how many tiles are chained into the (truncating) TMEM accumulators between drains."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath, RADIUS, SLOPE
C = gnnfc._cabi
w = WORKLOADS["cfg3"]; dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
st = hp.stream(); null = C.ct.c_void_p(0)
hp.fwd(0, st); torch.cuda.synchronize()
# fp64 reference of dH / db, in chunks
ref = torch.zeros(F, K, G, dtype=torch.float64, device=dev); refb = torch.zeros(F, dtype=torch.float64, device=dev)
CH = 2048
for b0 in range(0, B, CH):
    sl = slice(b0, b0 + CH)
    _, S = gnnfc.build_gso(hp.pos[0][sl], RADIUS, "binary_le")
    S = S.double()
    y = hp.y[0][sl].double(); D = hp.dY[0][sl].double()
    D = torch.where(y > 0, D, SLOPE * D)                       # [b,N,F]
    z = hp.x[0][sl].double()                                   # [b,G,N]
    refb += D.sum((0, 1))
    for k in range(K):
        ref[:, k, :] += torch.einsum("bnf,bgn->fg", D, z)
        z = torch.bmm(z, S)
ref = ref.reshape(-1)
def bwd():
    C.check(C.lib.gfc_filter_bwd_pos(C.ptr(hp.x[0]), C.ptr(hp.pos[0]), RADIUS, hp.mode, C.ptr(hp.h), C.ptr(hp.y[0]),
                                     C.ptr(hp.dY[0]), null, C.ptr(hp.dH), C.ptr(hp.db), B, N, G, F, K,
                                     C.ACT_LEAKY_RELU, SLOPE, C.PREC_FP32_3XTF32, C.ptr(hp.wsb), hp.nbb, st), "bwd")
for f in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 64]:
    C.check(C.lib.gfc_set_option(C.OPT_WIDE_FLUSH_EVERY, f), "opt")
    bwd(); torch.cuda.synchronize()
    err = ((hp.dH.double() - ref).abs().max() / ref.abs().max()).item()
    errb = ((hp.db.double() - refb).abs().max() / refb.abs().max()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): bwd()
    e1.record(); torch.cuda.synchronize()
    print("flush_every=%3d  dH err %.2e  db err %.2e  %.3f ms" % (f, err, errb, e0.elapsed_time(e1) / 3), flush=True)
