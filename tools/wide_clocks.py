#!/usr/bin/env python
"""Timeline of CTA 0 of the tcgen05 wide forward kernel (issuer thread + one worker warp)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gnnfc
from bench import WORKLOADS, HotPath
C = gnnfc._cabi
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]); w["B"] = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 6
dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 1)
st = hp.stream()
which = os.environ.get("WHICH", "fwd")
null = C.ct.c_void_p(0)
from bench import RADIUS, SLOPE
def run():
    if which == "fwd":
        hp.fwd(0, st)
    else:
        B, N, G, F, K = w["B"], w["N"], w["G"], w["F"], w["K"]
        dh = which in ("dh", "both")
        dx = which in ("bwd", "both")
        C.check(C.lib.gfc_filter_bwd_pos(C.ptr(hp.x[0]), C.ptr(hp.pos[0]), RADIUS, hp.mode, C.ptr(hp.h), C.ptr(hp.y[0]),
                                         C.ptr(hp.dY[0]), C.ptr(hp.dX[0]) if dx else null, C.ptr(hp.dH) if dh else null, C.ptr(hp.db) if dh else null, B, N, G, F, K, C.ACT_LEAKY_RELU, SLOPE,
                                         C.PREC_FP32_3XTF32, C.ptr(hp.wsb), hp.nbb, st), "bwd")
for _ in range(2): run()
torch.cuda.synchronize()
buf = torch.zeros(1184 * 16, dtype=torch.int64, device=dev)
C.check(C.lib.gfc_set_debug_clock_buffer(C.ptr(buf), buf.numel() * 8), "dbg")
run(); torch.cuda.synchronize()
C.lib.gfc_set_debug_clock_buffer(None, 0)
t = buf.cpu().numpy()
names = {100: "I wait w_ready s0", 101: "I wait w_ready s1", 110: "I got w_ready s0", 111: "I got w_ready s1", 120: "I hop issued s0", 121: "I hop issued s1",
         130: "I h_full 0", 131: "I h_full 1", 132: "I h_full 2", 133: "I h_full 3", 140: "I commit s0", 141: "I commit s1",
         200: "W wait mma_done s0", 201: "W wait mma_done s1", 210: "W got mma_done s0", 211: "W got mma_done s1", 220: "W stored s0", 221: "W stored s1",
         230: "W published s0", 231: "W published s1", 240: "W tail start", 241: "W P built", 250: "W last mma_done s0", 251: "W last mma_done s1",
         242: "W after worker_bar", 244: "W grp0 pairs", 245: "W grp0 st16", 246: "W grp1 pairs", 247: "W grp1 st16", 248: "W st_wait", 243: "W build_p done", 271: "W got out_full", 272: "W epi iter", 273: "W epi tmem loaded", 300: "I wait p_ready", 301: "I got p_ready", 310: "I wait v_ready", 311: "I got v_ready", 312: "I hop issued", 313: "I got x_ready", 314: "I dH issued", 400: "W wait hop_done", 401: "W got hop_done", 402: "W wrote back", 403: "W side jobs done", 404: "W loop done", 405: "W v0 loaded", 406: "W v0 stored", 407: "W item_done", 408: "W flushed", 409: "W xt stored", 260: "W w0 stored s0", 261: "W w0 stored s1", 270: "W epilogue done"}
ev = []
for base in (0, 2048):
    a = t[base:base + 2000].reshape(-1, 2)
    for c, tag in a:
        if c: ev.append((int(c), int(tag)))
ev.sort()
t0 = ev[0][0]
prev = t0
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
for c, tag in ev[lo:int(sys.argv[3]) if len(sys.argv) > 3 else 140]:
    print("%8d  +%6d  %s" % (c - t0, c - prev, names.get(tag, tag)))
    prev = c
