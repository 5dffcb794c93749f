#!/usr/bin/env python
"""torchrun -N ranks: whole-policy data-parallel step (encoder MLP -> graph filter -> action MLP, ~2.6 M parameters = the size
of the reference policy, suhaas_model.py) with gnnfc.BucketedReducer: step time without any exchange, with the collectives
launched after the backward (overlap=False) and with the per-bucket collectives launched from gradient hooks on a side stream
while the backward is still running (overlap=True).  Also checks that both modes produce the same averaged gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.distributed as dist
import gnnfc
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, N = int(os.environ.get("GFC_B", 2048)), 12

class Policy(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.Sequential(nn.Linear(1024, 1536), nn.ReLU(), nn.Linear(1536, 512), nn.ReLU(), nn.Linear(512, 128))
        self.gf = gnnfc.GraphFilterBatch(128, 128, 3, activation="leaky_relu")
        self.head = nn.Sequential(nn.Linear(128, 128), nn.LeakyReLU(), nn.Linear(128, 2))
    def forward(self, obs, pos):
        f = self.enc(obs).view(-1, N, 128).permute(0, 2, 1)          # [B,128,N] view over node-major memory
        self.gf.addPositions(pos, 2.0, "binary_le")
        return self.head(self.gf(f).permute(0, 2, 1))

torch.manual_seed(0)
model = Policy().to(dev)
gnnfc.broadcast_parameters(model.parameters())
params = list(model.parameters())
nparam = sum(p.numel() for p in params)
g = torch.Generator(device=dev).manual_seed(1 + rank)
obs = torch.randn(B * N, 1024, device=dev, generator=g)
pos = torch.rand(B, N, 2, device=dev, generator=g) * 6
tgt = torch.randn(B, N, 2, device=dev, generator=g)

def step(red):
    for p in params: p.grad = None
    loss = (model(obs, pos) - tgt).square().mean()
    loss.backward()
    if red is not None: red.finish()

def timeit(red, n=30):
    for _ in range(5): step(red)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(red)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

t_none = timeit(None)
res = {}
grads = {}
for overlap in (False, True):
    red = gnnfc.BucketedReducer(params, filter_params=list(model.gf.parameters()), bucket_bytes=4 << 20, average=True, overlap=overlap)
    res[overlap] = timeit(red)
    step(red); torch.cuda.synchronize()
    grads[overlap] = torch.cat([p.grad.flatten() for p in params]).clone()
    sizes = [b.numel * 4 for b in red.buckets]
    red.remove_hooks()
same = bool(torch.allclose(grads[False], grads[True], rtol=1e-5, atol=1e-6 * float(grads[False].abs().max())))
if world > 1:   # every rank ends with the same averaged gradients
    chk = [torch.empty_like(grads[True][:4096]) for _ in range(world)]
    dist.all_gather(chk, grads[True][:4096].contiguous())
    same &= all(torch.allclose(chk[0], c, rtol=1e-6, atol=1e-7) for c in chk)
if rank == 0:
    print("world %d, %d parameters (%.1f MB), buckets (bytes) %s, %d graphs x %d robots per rank" % (world, nparam, nparam * 4 / 1e6, sizes, B, N))
    print("step without exchange          %.3f ms" % t_none)
    print("exchange after the backward    %.3f ms  (+%.3f)" % (res[False], res[False] - t_none))
    print("exchange overlapped (hooks)    %.3f ms  (+%.3f)" % (res[True], res[True] - t_none))
    print("gradients identical across modes and ranks:", same, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
