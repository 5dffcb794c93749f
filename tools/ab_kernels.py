#!/usr/bin/env python
"""Per-kernel A/B of one gfc_set_option switch (bench.per_kernel_times: forward, dX-only call, whole backward call):
   tools/ab_kernels.py <cfg> <option key> <value A> <value B>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name, key, va, vb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
w = dict(bench.WORKLOADS[name])
dev = torch.device("cuda", 0)
import gnnfc
C = gnnfc._cabi
hp = bench.HotPath(w, dev, bench.ring_size(w))
for rnd in range(2):
    for v in (va, vb):
        C.check(C.lib.gfc_set_option(key, v), "gfc_set_option")
        for i in range(hp.ring): hp.step(i)
        kt = bench.per_kernel_times(torch, hp, 8)
        print("%s option %d = %d: " % (name, key, v) + ", ".join("%s %.3f ms" % (k, x["ms"]) for k, x in kt.items()), flush=True)
C.check(C.lib.gfc_set_option(key, va), "gfc_set_option")
