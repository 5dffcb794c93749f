#!/usr/bin/env python
"""Per-kernel timing of a workload's forward / backward-dX / backward-dH calls (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath, RADIUS, SLOPE
C = gnnfc._cabi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
flushes = [int(a) for a in sys.argv[2:]] or [0]
w = dict(WORKLOADS[name]); dev = torch.device("cuda", 0)
if os.environ.get("GFC_B"): w["B"] = int(os.environ["GFC_B"])
hp = HotPath(w, dev, 1)
if os.environ.get("NOPF"): C.check(C.lib.gfc_set_option(C.OPT_WIDE_NO_PREFETCH, 1), "opt")
B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
st = hp.stream()
null = C.ct.c_void_p(0)

def bwd(dx, dh):
    C.check(C.lib.gfc_filter_bwd_pos(C.ptr(hp.x[0]), C.ptr(hp.pos[0]), RADIUS, hp.mode, C.ptr(hp.h), C.ptr(hp.y[0]),
                                     C.ptr(hp.dY[0]), C.ptr(hp.dX[0]) if dx else null, C.ptr(hp.dH) if dh else null,
                                     C.ptr(hp.db) if dh else null, B, N, G, F, K, C.ACT_LEAKY_RELU, SLOPE,
                                     C.PREC_FP32_3XTF32, C.ptr(hp.wsb), hp.nbb, st), "bwd")

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

reps = 5 if B * N * G > 1e8 else 200
print(name, "fwd %.3f ms" % timeit(lambda: hp.fwd(0, st), reps), flush=True)
if w["train"]:
    print(name, "bwd dX only %.3f ms" % timeit(lambda: bwd(True, False), reps), flush=True)
    for f in flushes:
        if f: C.check(C.lib.gfc_set_option(C.OPT_WIDE_FLUSH_EVERY, f), "opt")
        print(name, "bwd dH+db only (flush_every=%d) %.3f ms" % (f, timeit(lambda: bwd(False, True), reps)), flush=True)
