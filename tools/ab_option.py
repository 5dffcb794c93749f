#!/usr/bin/env python
"""A/B of one gfc_set_option switch on a bench workload: whole step and backward call, CUDA-graph replay.
   tools/ab_option.py <cfg> <option key> <value A> <value B> [B]      e.g.  tools/ab_option.py cfg3 8 1 0"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name, key, va, vb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 5: w["B"] = int(sys.argv[5])
dev = torch.device("cuda", 0)
import gnnfc
C = gnnfc._cabi
ring = bench.ring_size(w)
hp = bench.HotPath(w, dev, ring)

def timed(fn, reps):
    for i in range(ring): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ring): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * ring)

reps = 10 if w["B"] * bench.bytes_per_graph(w)["total"] > 1e9 else 200
for rnd in range(2):
    for v in (va, vb):
        C.check(C.lib.gfc_set_option(key, v), "gfc_set_option")
        step = timed(lambda i: hp.step(i), reps)
        line = "%s option %d = %d: step %.4f ms" % (name, key, v, step)
        if w["train"]:
            line += ", backward call %.4f ms" % timed(lambda i: hp.bwd(i, hp.stream()), reps)
        print(line, flush=True)
C.check(C.lib.gfc_set_option(key, va), "gfc_set_option")
