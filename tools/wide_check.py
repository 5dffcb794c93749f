#!/usr/bin/env python
"""Quick parity check of the tcgen05 wide kernels against the oracle (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import gso as ogso, lsigf
import gnnfc
from test_gpu_parity import run_module, make_case, check_against_oracle
from util import rel_err

cases = {"cfg3_slice": (96, 64, 128, 128, 4, 10.0, 2), "cfg4_layer": (300, 12, 128, 128, 3, 6.0, 3),
         "cfg1": (64, 3, 128, 128, 3, 4.0, 0), "one_graph": (1, 64, 128, 128, 4, 10.0, 5),
         "k1": (40, 16, 128, 128, 1, 6.0, 6), "n128": (5, 128, 128, 128, 2, 12.0, 7),
         "c64": (77, 20, 64, 64, 3, 6.0, 8), "c64_128": (33, 9, 64, 128, 2, 4.0, 9), "big": (5000, 64, 128, 128, 4, 10.0, 10)}
names = sys.argv[1:] or list(cases)
if os.environ.get('GFC_FLUSH'):
    C = gnnfc._cabi
    C.check(C.lib.gfc_set_option(C.OPT_WIDE_FLUSH_EVERY, int(os.environ['GFC_FLUSH'])), 'opt')
for name in names:
    B, N, G, F, K, box, seed = cases[name]
    pos, h, b, x, dOut = make_case(B, N, G, F, K, box, seed)
    S64, _ = ogso.gso(pos, 2.0, ogso.MODE_BINARY_LE)
    got = run_module(h, b, x, dOut, pos=pos, radius=2.0, mode="binary_le", act="leaky_relu")
    try:
        errs = check_against_oracle(h, b, S64[:, None], x, dOut, got, lsigf.ACT_LEAKY_RELU, tol=1.0, tag=name)
        print(name, "y %.2e dX %.2e dH %.2e db %.2e" % tuple(errs), flush=True)
    except AssertionError as ex:
        print(name, "FAILED", str(ex)[:200], flush=True)
