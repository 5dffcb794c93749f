"""CPU, world_size 2 over gloo: the data-parallel plumbing (contiguous batch
sharding + one flat-bucket all-reduce of [dH | db]) reproduces the single-process
gradient of the concatenated batch.  The per-rank gradient itself comes from the
oracle here (no GPU in this container); on the GPU box the same code path is fed
by kernel (c) — see tests/test_gpu_parity.py::test_dp_two_shards_equal_full_batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnnfc
from oracle import lsigf


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _case():
    rng = np.random.default_rng(42)
    B, N, G, F, K = 10, 5, 4, 6, 3
    h = rng.standard_normal((F, 1, K, G)).astype(np.float32)
    b = rng.standard_normal((F, 1)).astype(np.float32)
    S = (rng.random((B, 1, N, N)) < 0.4).astype(np.float32)
    x = rng.standard_normal((B, G, N)).astype(np.float32)
    dY = rng.standard_normal((B, F, N)).astype(np.float32)
    return h, b, S, x, dY


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, b, S, x, dY = _case()
    lo, hi = gnnfc.shard_range(x.shape[0], rank, world)
    w = torch.nn.Parameter(torch.from_numpy(h.copy() + (rank * 0.5)))   # deliberately different ...
    bb = torch.nn.Parameter(torch.from_numpy(b.copy()))
    gnnfc.broadcast_parameters([w, bb], src=0)                           # ... until broadcast
    _, dH, db = lsigf.lsigf_backward(w.detach().numpy(), S[lo:hi], x[lo:hi], dY[lo:hi])
    w.grad = torch.from_numpy(dH).float(); bb.grad = torch.from_numpy(db).float()
    bucket = gnnfc.GradBucket([w, bb], average=False)
    bucket.sync_grads()
    if rank == 0:
        out.put((w.detach().numpy(), w.grad.numpy(), bb.grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    w, gw, gb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    h, b, S, x, dY = _case()
    assert np.array_equal(w, h)                                         # broadcast took rank 0's taps
    _, dH, db = lsigf.lsigf_backward(h, S, x, dY)
    assert np.allclose(gw, dH, rtol=1e-5, atol=1e-5 * np.abs(dH).max())
    assert np.allclose(gb, db, rtol=1e-5, atol=1e-5 * np.abs(db).max())
