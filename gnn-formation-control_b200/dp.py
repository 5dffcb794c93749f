"""Data-parallel plumbing for the graph-filter path.

Graphs are independent units: the batch is sharded contiguously over ranks, the
forward / backward kernels run rank-local with no exchange, and one flat fp32
bucket ``[dH | db | (other grads)]`` is all-reduced per step.  The reference has
no distributed code at all (single ``.to('cuda')``, suhaas_agent.py:19); this is
the layer the north star adds.  Backend: NCCL over NVLink on GPUs, gloo in the
CPU unit tests (the bucketing logic is backend-agnostic).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """contiguous shard [lo, hi) of ``total`` graphs for ``rank``; remainders go to
    the lowest ranks so sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat fp32 bucket over a fixed parameter list; one collective per step."""

    def __init__(self, params, average=True, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.average, self.group = average, process_group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def pack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        return self.flat

    def unpack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n

    def allreduce(self):
        """sum (or mean) of the bucket over all ranks; async_op-free, stream-ordered."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat

    def sync_grads(self):
        self.pack()
        self.allreduce()
        self.unpack()


def broadcast_parameters(params, src=0, group=None):
    """identical taps on every rank before the first step"""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
