#!/usr/bin/env python
"""Phase breakdown of the fused tile kernels (SM clock stamps of every CTA's first tile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gnnfc
from bench import WORKLOADS, HotPath, ring_size
C = gnnfc._cabi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = WORKLOADS[name]; dev = torch.device("cuda", 0)
hp = HotPath(w, dev, 2)
buf = torch.zeros(1184 * 16, dtype=torch.int64, device=dev)
# stamps of thread 0 (warp 0): 0 = first loads issued, 1 = produce(tile 0) done, 2 = produce(tile 1) done,
# 3 = first epilogue (fwd: y stored; bwd: dX stored), 4 = bwd: dH tile drained, 6 = bwd: partials written
for which, labels in (("fwd", ["setup", "produce(0)", "produce(1)", "epilogue(0)"]),
                      ("bwd", ["setup", "produce(0)", "produce(1)", "dX epilogue(0)", "dH epilogue(0)"])):
    if which == "bwd" and not w["train"]:
        continue
    for _ in range(3):
        hp.step(0)
    torch.cuda.synchronize()
    buf.zero_()
    C.check(C.lib.gfc_set_debug_clock_buffer(C.ptr(buf), buf.numel() * 8), "dbg")
    st = hp.stream()
    (hp.fwd if which == "fwd" else hp.bwd)(1, st)
    torch.cuda.synchronize()
    C.lib.gfc_set_debug_clock_buffer(None, 0)
    t = buf.cpu().numpy().reshape(-1, 16)
    t = t[t[:, 0] != 0]
    n = len(labels)
    d = np.diff(t[:, :n], axis=1).astype(np.float64)
    print("%s %s: %d CTAs; per-phase cycles median [p10, p90]; total median %.0f" % (name, which, len(t), np.median(t[:, n - 1] - t[:, 0])))
    for i in range(n - 1):
        col = d[:, i]
        print("   %-12s %8.0f [%6.0f, %6.0f]" % (labels[i + 1], np.median(col), np.percentile(col, 10), np.percentile(col, 90)))
    print("   setup (entry -> stamp 0): %.0f cycles median" % np.median(t[:, 0] - t[:, 7]))
    ns0, ns1 = t[:, 8], t[:, 9]
    print("   globaltimer: CTA starts spread %.2f us, CTA duration median %.2f us, first start -> last end %.2f us"
          % ((ns0.max() - ns0.min()) / 1e3, np.median(ns1 - ns0) / 1e3, (ns1.max() - ns0.min()) / 1e3))
