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
    """Flat fp32 bucket over a fixed parameter list; one collective per step.

    On GPUs (NCCL process group) buckets of up to ``PEER_MAX_NUMEL`` floats are all-reduced by the one-shot
    peer-memory kernel (``PeerExchange``: CUDA symmetric memory over NVLink, created collectively on the first
    call — every rank must reach its first ``allreduce`` together); larger buckets, CPU tensors and other backends
    use ``torch.distributed.all_reduce``.  Set ``PEER_MAX_NUMEL = 0`` on the instance to force the latter."""

    # buckets up to this many floats go through the one-shot peer-memory kernel (latency-bound regime: every rank
    # pushes its whole bucket to every peer); larger ones through NCCL (bandwidth-optimal rings / NVLS)
    PEER_MAX_NUMEL = 1 << 16

    def __init__(self, params, average=True, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.average, self.group = average, process_group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self._px = None          # PeerExchange, created on the first GPU all-reduce (collective: all ranks together)

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

    def _agree_on_peer_exchange(self, world):
        """Collective decision, taken once: EVERY rank builds the peer exchange or none uses it.  Construction can
        fail on some ranks only (no symmetric memory in the torch build, no P2P / NVLink between two devices, IPC
        disabled, a multi-node group, world > 16); a rank that silently fell back to NCCL while its peers entered the
        one-shot kernel would leave them waiting for words that never come.  So each rank tries, then the ranks
        all-reduce an ok flag with MIN and follow the common verdict."""
        px, err = None, None
        if world <= PeerExchange.MAX_WORLD:
            try:
                px = PeerExchange(self.numel, self.flat.device, self.group)
            except Exception as ex:   # noqa: BLE001 — any failure means "not on this rank", decided collectively below
                err = ex
        ok = torch.tensor([1 if px is not None else 0], dtype=torch.int32, device=self.flat.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            return px
        self.peer_error = err    # kept for diagnosis; the NCCL path is used by every rank
        return False

    def allreduce(self):
        """sum (or mean) of the bucket over all ranks; async_op-free, stream-ordered."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return self.flat
        world = dist.get_world_size(self.group)
        if (self._px is not False and self.flat.is_cuda and 0 < self.numel <= self.PEER_MAX_NUMEL
                and dist.get_backend(self.group) == "nccl"):
            if self._px is None:
                self._px = self._agree_on_peer_exchange(world)
            if self._px:
                self._px.allreduce_(self.flat, scale=(1.0 / world) if self.average else 1.0)
                return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            self.flat.div_(world)
        return self.flat

    def sync_grads(self):
        self.pack()
        self.allreduce()
        self.unpack()


class BucketedReducer:
    """Data-parallel gradient exchange for the WHOLE policy — filter taps and encoder / MLP weights — overlapped with
    the backward pass (the reference trains on one device, suhaas_agent.py:115-128).

    Parameters are grouped into buckets in reverse registration order (the order autograd finishes them, as the
    reference model is wired: action MLP -> graph filter -> compress MLP -> CNN, suhaas_model.py:53-143,161-212).
    ``filter_params`` (the ``GraphFilterBatch`` taps / biases) get a bucket of their own, small enough for the one-shot
    peer-memory kernel; everything else is cut into ``bucket_bytes`` (default 10 MB) buckets for NCCL.  Every
    parameter carries a post-accumulate-grad hook: the gradient is copied into its bucket slice as soon as autograd
    has produced it, and when a bucket is complete its all-reduce is launched on a SIDE stream (ordered after the
    copies through an event) while the main stream continues with the rest of the backward — e.g. the action-MLP
    bucket travels while the filter backward kernels run, the filter bucket while the CNN backward runs.
    ``finish()`` makes the main stream wait for every bucket and writes the reduced values back into ``p.grad``.

    ``overlap=False`` launches the same collectives on the main stream inside ``finish()`` (the A/B baseline).
    CPU tensors / gloo: same bucketing and hooks, collectives run inline (tests/test_dp_gloo.py)."""

    def __init__(self, params, filter_params=(), bucket_bytes=10 << 20, average=True, process_group=None,
                 overlap=True):
        params = [p for p in params if p.requires_grad]
        fset = {id(p) for p in filter_params}
        self.average, self.group, self.overlap = average, process_group, overlap
        groups, cur, cur_bytes = [], [], 0
        for p in reversed([p for p in params if id(p) not in fset]):
            nb = p.numel() * 4
            if cur and cur_bytes + nb > bucket_bytes:
                groups.append(cur); cur, cur_bytes = [], 0
            cur.append(p); cur_bytes += nb
        if cur:
            groups.append(cur)
        fl = [p for p in params if id(p) in fset]
        if fl:
            groups.append(fl)
        self.buckets = [GradBucket(g, average=average, process_group=process_group) for g in groups]
        self.filter_bucket = len(self.buckets) - 1 if fl else None
        self._slot = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b.params:
                self._slot[id(p)] = (bi, off, p.numel())
                off += p.numel()
        self._pending = [len(b.params) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._done = [None] * len(self.buckets)
        dev = params[0].device if params else torch.device("cpu")
        self.cuda = dev.type == "cuda"
        self.side = torch.cuda.Stream(device=dev) if (self.cuda and overlap) else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.launch_order = []      # bucket indices in the order their collectives were launched (last step)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def total_numel(self):
        return sum(b.numel for b in self.buckets)

    def _on_grad(self, p):
        bi, off, n = self._slot[id(p)]
        b = self.buckets[bi]
        b.flat[off:off + n].copy_(p.grad.reshape(-1))
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.overlap:
            self._launch(bi)

    def _launch(self, bi):
        b = self.buckets[bi]
        self._launched[bi] = True
        self.launch_order.append(bi)
        if self.side is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(b.flat.device))
            self.side.wait_event(ready)
            with torch.cuda.stream(self.side):
                b.allreduce()
                done = torch.cuda.Event()
                done.record(self.side)
            self._done[bi] = done
        else:
            b.allreduce()

    def finish(self):
        """wait for / run the outstanding collectives, then unpack into ``p.grad``; call after ``loss.backward()``"""
        for bi, b in enumerate(self.buckets):
            if not self._launched[bi]:
                if self._pending[bi]:              # parameters that received no gradient this step contribute zeros
                    for p in b.params:
                        if p.grad is None:
                            _, off, n = self._slot[id(p)]
                            b.flat[off:off + n].zero_()
                self._launch(bi)
        for bi, b in enumerate(self.buckets):
            if self._done[bi] is not None:
                torch.cuda.current_stream(b.flat.device).wait_event(self._done[bi])
                self._done[bi] = None
            b.unpack()
        self._pending = [len(b.params) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        order, self.launch_order = self.launch_order, []
        return order


def broadcast_parameters(params, src=0, group=None):
    """identical taps on every rank before the first step"""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)


class PeerExchange:
    """Peer-mapped exchange + signal buffers for the fused gradient reduction / all-reduce kernel
    (``gfc_filter_bwd*_dp``, ``gfc_dp_allreduce``; csrc/gfc_dp.cu).

    The buffers live in CUDA symmetric memory (``torch.distributed._symmetric_memory``): every rank allocates the
    same sizes, the rendezvous maps all of them into every process, and the kernels store into / poll them
    directly over NVLink.  ``world == 1`` (or no process group) needs no mapping: the rank talks to itself."""

    MAX_WORLD = 16   # GFC_DP_MAX_WORLD (csrc/gfc_dp.cuh)

    def __init__(self, numel, device, group=None):
        import ctypes as ct
        from . import _cabi as C
        self.C, self.n = C, int(numel)
        init = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if init else 0
        self.world = dist.get_world_size(group) if init else 1
        nb = C.lib.gfc_dp_exchange_bytes(self.n, self.world)
        ns = C.lib.gfc_dp_signal_bytes(self.n, self.world)
        if not (nb > 0 and ns > 0):
            raise ValueError("PeerExchange: bucket of %d floats / world size %d not supported (world <= %d)"
                             % (self.n, self.world, self.MAX_WORLD))
        self.bytes = nb + ns
        self._sig_off = nb
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.mem = symm.empty(self.bytes, dtype=torch.uint8, device=device)
            self.mem.zero_()
            torch.cuda.synchronize(device)
            self.handle = symm.rendezvous(self.mem, group=(group or dist.group.WORLD))
            bases = [int(p) for p in self.handle.buffer_ptrs]
            dist.barrier(group)
        else:
            self.mem = torch.zeros(self.bytes, dtype=torch.uint8, device=device)
            self.handle = None
            bases = [self.mem.data_ptr()]
        arr = ct.c_void_p * self.world
        self.buf_ptrs = arr(*[ct.c_void_p(b) for b in bases])
        self.sig_ptrs = arr(*[ct.c_void_p(b + nb) for b in bases])

    def allreduce_(self, flat, scale=1.0, stream=None):
        """in-place all-reduce of a flat fp32 CUDA tensor (sum * scale) through the fused kernel"""
        C = self.C
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() == self.n
        st = C.ct.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream if stream is None else stream)
        C.check(C.lib.gfc_dp_allreduce(C.ptr(flat), C.ptr(flat), self.n, self.buf_ptrs, self.sig_ptrs,
                                       self.rank, self.world, float(scale), st), "gfc_dp_allreduce")
        return flat

    def status(self):
        """SYNCHRONISING health check: raises GfcError (GFC_ERR_TIMEOUT) if any launch of the exchange gave up waiting
        for a peer (the affected gradient elements are then NaN).  The kernel itself never hangs (bounded poll,
        ``GFC_OPT_DP_TIMEOUT_MS``)."""
        C = self.C
        sig = C.ct.c_void_p(self.mem.data_ptr() + self._sig_off)
        missing = C.ct.c_int(-1)
        st = C.ct.c_void_p(torch.cuda.current_stream(self.mem.device).cuda_stream)
        with torch.cuda.device(self.mem.device):
            C.check(C.lib.gfc_dp_status(sig, self.n, C.ct.byref(missing), st), "gfc_dp_status")
        return True
