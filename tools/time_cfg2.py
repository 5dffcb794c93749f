#!/usr/bin/env python
"""cfg2 step / kernel timings with programmatic dependent launch on and off (graph replay, rotating batches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnfc
from bench import WORKLOADS, HotPath, ring_size, timed_steps, kernel_alone_ms
C = gnnfc._cabi
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = dict(WORKLOADS[name]); dev = torch.device("cuda", 0)
if len(sys.argv) > 2:
    w["B"] = int(sys.argv[2])
hp = HotPath(w, dev, ring_size(w))
for pdl in (1, 0, 1):
    C.check(C.lib.gfc_set_option(C.OPT_PDL, pdl), "opt")
    steps = 2000 if w["B"] <= 8192 else 64
    steps -= steps % hp.ring
    ms = timed_steps(torch, hp, steps, 5, 1, None, True)
    f, _ = kernel_alone_ms(torch, hp, "fwd", 40)
    b, _ = kernel_alone_ms(torch, hp, "bwd", 40)
    print("%s B=%d pdl=%d: step %.2f us (%.1f M graphs/s), fwd alone %.2f us, bwd alone %.2f us" %
          (name, w["B"], pdl, ms / steps * 1e3, w["B"] * steps / ms / 1e3, f * 1e3, b * 1e3))
