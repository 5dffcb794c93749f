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


# --------------------------------------------------------------------------- #
# §8 f-1: whole-policy gradient exchange — bucketing, hooks, launch order — with the reference policy's own
# parameter layout (2 584 034 parameters, suhaas_model.py:53-143; tests/golden/model_param_layout.json was written
# from the constructed reference model by oracle/make_golden.py)
# --------------------------------------------------------------------------- #
def _policy_params(seed):
    import json
    lay = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_param_layout.json")))
    g = torch.Generator().manual_seed(seed)
    return [(n, torch.nn.Parameter(torch.randn(*s, generator=g))) for n, s in lay]


def _reducer_worker(rank, world, port, out, overlap):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    named = _policy_params(0)
    params = [p for _, p in named]
    filt = [p for n, p in named if n.startswith("GFL.")]
    red = gnnfc.BucketedReducer(params, filter_params=filt, bucket_bytes=4 << 20, average=True, overlap=overlap)
    # a loss whose gradient differs per rank and per parameter: d/dp sum(c_rank_i * p) = c_rank_i
    res = []
    for step in range(2):
        for p in params:
            p.grad = None
        loss = sum(((rank + 1) * (i + 1 + step)) * p.sum() for i, p in enumerate(params))
        loss.backward()
        order = red.finish()
        res.append((order, [float(p.grad.flatten()[0]) for p in params],
                    all(bool((p.grad == p.grad.flatten()[0]).all()) for p in params)))
    if rank == 0:
        out.put((res, [b.numel for b in red.buckets], red.filter_bucket, red.total_numel,
                 [[next(n for n, p in named if p is q) for q in b.params] for b in red.buckets]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_with_the_reference_policy_layout():
    ctx = mp.get_context("spawn")
    for overlap in (True, False):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
        for p in procs:
            p.start()
        res, sizes, fb, total, names = q.get(timeout=300)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        assert total == 2584034 and sum(sizes) == total
        assert fb == len(sizes) - 1 and names[fb] == ["GFL.0.weight", "GFL.0.bias"] and sizes[fb] == 128 * 3 * 128 + 128
        assert all(s * 4 <= (4 << 20) or len(n) == 1 for s, n in zip(sizes, names))      # cut at 4 MB here
        assert names[0][0].startswith("actionsMLP")                                   # reverse registration order
        nparam = sum(len(n) for n in names)
        for step, (order, first, uniform) in enumerate(res):
            assert sorted(order) == list(range(len(sizes)))                          # every bucket exactly once
            if overlap:     # launched as autograd completes them; finish() only waits
                assert len(order) == len(sizes)
            assert uniform
            # mean over ranks of (rank+1)*(i+1+step) = 1.5*(i+1+step)
            assert np.allclose(first, [1.5 * (i + 1 + step) for i in range(nparam)], rtol=1e-6)
