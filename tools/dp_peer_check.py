#!/usr/bin/env python
"""torchrun -N ranks: fused reduce + one-shot all-reduce (gfc_dp.cu) vs NCCL, and DP gradients vs the full batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gnnfc
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 3104
px = gnnfc.PeerExchange(n, dev)
ok = True
for it in range(20):
    g = torch.Generator(device=dev).manual_seed(100 * it + rank)
    v = torch.randn(n, device=dev, generator=g)
    ref = v.clone(); dist.all_reduce(ref)
    got = px.allreduce_(v.clone())
    torch.cuda.synchronize()
    err = float((got - ref).abs().max() / ref.abs().max())
    ok &= err < 1e-6
    allv = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(allv, got)
    ok &= all(torch.equal(allv[0], a) for a in allv)     # bit-identical on every rank
if rank == 0: print("peer all-reduce vs NCCL ok:", ok, "err %.2e" % err, flush=True)
# GradBucket over a module's parameters: small buckets take the fused peer kernel, averages included
lin = torch.nn.Linear(24, 17).to(dev)
bucket = gnnfc.GradBucket(lin.parameters(), average=True)
for p_ in lin.parameters():
    p_.grad = torch.full_like(p_, float(rank + 1))
bucket.sync_grads()
torch.cuda.synchronize()
want = sum(range(1, world + 1)) / world
okb = all(bool(torch.allclose(p_.grad, torch.full_like(p_, want))) for p_ in lin.parameters()) and bool(bucket._px) and px.status() and bucket._px.status()
if rank == 0: print("GradBucket through the peer exchange ok:", okb, flush=True)
# timing
for name, fn in (("fused peer exchange", lambda t: px.allreduce_(t)), ("NCCL", lambda t: dist.all_reduce(t))):
    t = torch.randn(n, device=dev)
    for _ in range(20): fn(t)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn(t)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("%-22s %.2f us per all-reduce of %d floats at world %d" % (name, e0.elapsed_time(e1) * 5, n, world), flush=True)
dist.barrier(); dist.destroy_process_group()
