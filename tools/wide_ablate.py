#!/usr/bin/env python
"""Ablation timing of the tcgen05 wide forward kernel: GFC_OPT_WIDE_NO_PREFETCH bit mask
(1 no L2 prefetch, 2 skip epilogue, 4 skip P build, 8 skip write-back work) -> ms per launch. Results are WRONG
with bits 2/4/8 set; this only shows where the per-tile time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath
C = gnnfc._cabi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = WORKLOADS[name]; dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
st = hp.stream()
for mask in (0, 16, 14, 30, 0):
    C.check(C.lib.gfc_set_option(C.OPT_WIDE_NO_PREFETCH, mask), "opt")
    for _ in range(2): hp.fwd(0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): hp.fwd(0, st)
    e1.record(); torch.cuda.synchronize()
    print(name, "mask", mask, "fwd %.3f ms" % (e0.elapsed_time(e1) / 5), flush=True)
C.check(C.lib.gfc_set_option(C.OPT_WIDE_NO_PREFETCH, 0), "opt")
