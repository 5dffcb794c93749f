#!/usr/bin/env python
"""Timelines of CTA 0 of the tcgen05 wide kernels (needs a library built with EXTRA=-DGFC_WIDE_TIMELINE): SM-clock stamps
(gfc_set_debug_clock_buffer) of the issuing thread (I), the first write-back warp (W) and — forward / dX kernel — the first
warp of group B.   usage: wide_clocks.py [cfg] [B] [fwd|dx|dh] [first_event] [n_events]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gnnfc
from bench import WORKLOADS, HotPath, RADIUS, SLOPE
C = gnnfc._cabi
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]); w["B"] = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 8
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
st = hp.stream()
B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
null = C.ct.c_void_p(0)
def run():
    if which == "fwd":
        hp.fwd(0, st); return
    dx, dh = which == "dx", which == "dh"
    C.check(C.lib.gfc_filter_bwd_pos(C.ptr(hp.x[0]), C.ptr(hp.pos[0]), RADIUS, hp.mode, C.ptr(hp.h), C.ptr(hp.y[0]),
                                     C.ptr(hp.dY[0]), C.ptr(hp.dX[0]) if dx else null, C.ptr(hp.dH) if dh else null,
                                     C.ptr(hp.db) if dh else null, B, N, G, F, K, C.ACT_LEAKY_RELU, SLOPE,
                                     C.PREC_FP32_3XTF32, C.ptr(hp.wsb), hp.nbb, st), "bwd")
hp.fwd(0, st)
for _ in range(2): run()
torch.cuda.synchronize()
buf = torch.zeros(1184 * 16, dtype=torch.int64, device=dev)
C.check(C.lib.gfc_set_debug_clock_buffer(C.ptr(buf), buf.numel() * 8), "dbg")
run(); torch.cuda.synchronize()
C.lib.gfc_set_debug_clock_buffer(None, 0)
t = buf.cpu().numpy()
def name(tag):
    if tag == 1: return "I tile start"
    if tag == 3: return "I got p_ready"
    if 100 <= tag < 200: return "I   wait operand ph%d" % (tag - 100)
    if 200 <= tag < 250: return "I   got  operand ph%d" % (tag - 200)
    if 250 <= tag < 260: return "I   hop issued k%d" % (tag - 250)
    if 260 <= tag < 270: return "I   x ready k%d" % (tag - 260)
    if 300 <= tag < 400: return "I   issued+commit ph%d" % (tag - 300)
    return "I %d" % tag
def wname(tag):
    if 100 <= tag < 200: return "W wait hop_done %d" % (tag - 100)
    if 200 <= tag < 300: return "W got  hop_done %d" % (tag - 200)
    if 300 <= tag < 400: return "W wrote back + published %d" % (tag - 300)
    if 400 <= tag < 410 and which == "dh": return "W gap work done %d" % (tag - 400)
    return {401: "inputs loaded, tile max done", 430: "P published", 410: "slab0 loaded+free (dh: before store_v0)", 411: "slab1 loaded+free",
            420: "W0 slab0 stored (dh: V0 published)", 421: "W0 slab1 stored", 500: "wait out_full/item_done", 501: "epilogue done (dh: got item_done)",
            502: "dh: flushed", 510: "dh: X^T stored"}.get(tag, "%d" % tag)
ev = []
a = t[0:4000].reshape(-1, 2)
ev += [(int(c), name(int(tag))) for c, tag in a if c]
a = t[4096:4096 + 4000].reshape(-1, 2)
ev += [(int(c), "        " + wname(int(tag))) for c, tag in a if c]
a = t[8192:8192 + 4000].reshape(-1, 2)     # group B (next tile's operands, previous tile's epilogue)
ev += [(int(c), "                        B " + wname(int(tag))) for c, tag in a if c]
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 150
n = int(sys.argv[5]) if len(sys.argv) > 5 else 110
prev = ev[lo][0]
print("== %s %s B=%d: events %d..%d of %d" % (sys.argv[1] if len(sys.argv) > 1 else "cfg3", which, B, lo, lo + n, len(ev)))
for c, nm in ev[lo:lo + n]:
    print("%8d  +%6d  %s" % (c - t0, c - prev, nm))
    prev = c
