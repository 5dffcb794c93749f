#!/usr/bin/env python
"""Timeline of CTA 0 of the tcgen05 wide forward kernel (issuer thread, first warp of worker group A (W) and of group B): SM-clock stamps
(gfc_set_debug_clock_buffer).  usage: wide_clocks.py [cfg] [B] [first_event] [n_events]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gnnfc
from bench import WORKLOADS, HotPath
C = gnnfc._cabi
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]); w["B"] = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 8
dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
st = hp.stream()
if os.environ.get("MASK"): C.check(C.lib.gfc_set_option(C.OPT_WIDE_NO_PREFETCH, int(os.environ["MASK"])), "opt")
for _ in range(2): hp.fwd(0, st)
torch.cuda.synchronize()
buf = torch.zeros(1184 * 16, dtype=torch.int64, device=dev)
C.check(C.lib.gfc_set_debug_clock_buffer(C.ptr(buf), buf.numel() * 8), "dbg")
hp.fwd(0, st); torch.cuda.synchronize()
C.lib.gfc_set_debug_clock_buffer(None, 0)
t = buf.cpu().numpy()
def name(tag):
    if tag == 1: return "I tile start"
    if tag == 2: return "I got out_free"
    if tag == 3: return "I got p_ready"
    if 100 <= tag < 200: return "I   wait w_ready ph%d" % (tag - 100)
    if 200 <= tag < 300: return "I   got  w_ready ph%d" % (tag - 200)
    if 300 <= tag < 400: return "I   issued+commit ph%d" % (tag - 300)
    return str(tag)
def wname(tag):
    if 100 <= tag < 200: return "        W wait mma_done ph%d" % (tag - 100)
    if 200 <= tag < 300: return "        W got  mma_done ph%d" % (tag - 200)
    if 300 <= tag < 400: return "        W wrote back + published ph%d" % (tag - 300)
    return {400: "        W write-backs done", 401: "        W inputs loaded, tile max done", 430: "        W P published", 410: "        W slab0 free", 411: "        W slab1 free",
            420: "        W W0 slab0 stored", 421: "        W W0 slab1 stored", 500: "        W got out_full", 501: "        W epilogue done"}.get(tag, "        W %d" % tag)
ev = []
a = t[0:4000].reshape(-1, 2)
ev += [(int(c), name(int(tag))) for c, tag in a if c]
a = t[4096:4096 + 4000].reshape(-1, 2)
ev += [(int(c), wname(int(tag))) for c, tag in a if c]
a = t[8192:8192 + 4000].reshape(-1, 2)     # group B (next tile's operands, previous tile's epilogue)
ev += [(int(c), "                        B" + wname(int(tag)).strip().lstrip("W")) for c, tag in a if c]
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 150
n = int(sys.argv[4]) if len(sys.argv) > 4 else 110
prev = ev[lo][0]
for c, nm in ev[lo:lo + n]:
    print("%8d  +%6d  %s" % (c - t0, c - prev, nm))
    prev = c
