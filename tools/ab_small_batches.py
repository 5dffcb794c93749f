#!/usr/bin/env python
"""A/B of the two kernel families at SMALL batches (few 128-row tiles): tcgen05 wide path vs the mma.sync tile kernels
(GFC_OPT_DISABLE_TCGEN05), step time by CUDA-graph replay.   tools/ab_small_batches.py [cfg1|cfg4] [B ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
Bs = [int(a) for a in sys.argv[2:]] or [bench.WORKLOADS[name]["B"]]
dev = torch.device("cuda", 0)
import gnnfc
C = gnnfc._cabi
for B in Bs:
    w = dict(bench.WORKLOADS[name]); w["B"] = B
    for off in (0, 1):
        C.check(C.lib.gfc_set_option(C.OPT_DISABLE_TCGEN05, off), "gfc_set_option")
        hp = bench.HotPath(w, dev, 8)
        for i in range(8): hp.step(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(8): hp.step(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("%s B=%d tiles=%d %s: %.2f us/step, %d launches/step, path %d" % (
            name, B, -(-B // max(1, 128 // w["N"])), "mma.sync tile kernels" if off else "tcgen05 wide kernels  ",
            e0.elapsed_time(e1) / 400 * 1e3, hp.launches_per_step, C.last_path()))
        del hp, g
C.check(C.lib.gfc_set_option(C.OPT_DISABLE_TCGEN05, 0), "gfc_set_option")
