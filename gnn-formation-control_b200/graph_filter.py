"""Drop-in ``GraphFilterBatch`` backed by the sm_100a kernels in libgfc.so.

Mirrors the reference layer one-to-one (utils/graphUtils/graphML.py:2369-2488):
same constructor ``GraphFilterBatch(G, F, K, E=1, bias=True)``, same parameter
names and shapes (``weight [F,E,K,G]``, ``bias [F,1]`` — so reference
``state_dict`` keys ``...GFL.0.weight/bias`` load unchanged), same
``reset_parameters`` law, same ``addGSO`` asserts, same ``forward`` semantics
(zero-pad when ``Nin < N``, slice back), same ``extra_repr`` string, and the
output is the same strided ``[B,F,N]`` view over ``[B,N,F]`` memory.

Additive API (not in the reference): ``addPositions`` (the GSO is rebuilt on
chip from robot positions — scene.py:140-154 / multirobotsim_dcenlocal.py:291-317
— and never exists in HBM), ``addSparseGSO`` (CSR path for large swarms), a
fused activation, and the arithmetic mode of the tap contraction.

dtype: the reference computes in float64 (graphML.py:2350,2361-2362).  The
kernels compute in float32 (tap contraction = 3xTF32 split on tensor cores,
fp32-equivalent); inputs of any float dtype are accepted and cast; the result is
float32, or float64 with ``reference_dtype=True``.
"""
import math

import torch
import torch.nn as nn

from . import _cabi as C
from .gso import SparseGSO, build_csr

_SRC_DENSE, _SRC_POS, _SRC_CSR = 0, 1, 2


def _stream():
    return C.ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes, device):
    if nbytes == 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("gnnfc.GraphFilterBatch: %s must be a CUDA tensor — this layer has no CPU "
                           "fallback (the reference's CPU path lives in oracle/ for tests only)" % what)


def _is_binary(S32, E, G, F):
    """True when a dense GSO holds only 0/1 entries (what Scene.readADjMatrix produces, scene.py:140-154) AND the layer
    has a shape the tcgen05 wide kernels serve — only then is the check (one reduction + one host read per addGSO)
    worth making.  During CUDA-graph capture the host read is impossible: the generic dense kernels are used."""
    if E != 1 or G not in (64, 128) or F not in (64, 128) or S32.shape[-1] > 127:
        return False
    if torch.cuda.is_current_stream_capturing():
        return False
    return bool(((S32 == 0) | (S32 == 1)).all().item())


class _Src:
    """graph source handed to the autograd function (not a tensor argument)."""

    def __init__(self, kind, S=None, pos=None, radius=0.0, mode=0, csr=None, binary=False, shared=False):
        self.kind, self.S, self.pos, self.radius, self.mode, self.csr = kind, S, pos, radius, mode, csr
        self.binary = binary     # dense S with 0/1 entries only: eligible for the tcgen05 wide kernels
        self.shared = shared     # dense S is [E,N,N]: one GSO for the whole batch (GraphFilter)

    def flags(self):
        return (C.PREC_FLAG_BINARY_GSO if self.binary else 0) | (C.PREC_FLAG_SHARED_GSO if self.shared else 0)


class _LSIGF(torch.autograd.Function):
    """y_mem[B,N,F] = act(sum_k z_k H_k + b);  backward = kernel (c)."""

    @staticmethod
    def forward(ctx, x, weight, bias, src, act, slope, prec):
        F_, E, K, G = weight.shape
        B, Gx, N = x.shape
        assert Gx == G
        dev = x.device
        w32 = weight.detach().to(torch.float32).contiguous()
        b32 = bias.detach().to(torch.float32).contiguous().view(-1) if bias is not None else None
        y = torch.empty((B, N, F_), dtype=torch.float32, device=dev)
        # Stacked layers: the previous layer's output is a [B,G,N] VIEW over node-major [B,N,G] memory (the
        # reference's own layout, graphML.py:2362).  Where the node-major entry point applies it is consumed in
        # place — no transposing copy between the layers.
        x32 = None
        if (src.kind == _SRC_POS and x.dtype == torch.float32 and N > 1 and G > 1
                and x.stride() == (N * G, 1, G)):
            with torch.cuda.device(dev):
                nb = C.lib.gfc_filter_workspace_bytes(B, N, G, F_, K, 1, 0)
                ws = _workspace(nb, dev)
                rc = C.lib.gfc_filter_fwd_pos_nm(C.ptr(x), C.ptr(src.pos), src.radius, src.mode, C.ptr(w32),
                                                 C.ptr(b32), C.ptr(y), B, N, G, F_, K, act, slope, prec,
                                                 C.ptr(ws), nb, _stream())
            if rc == C.GFC_OK:
                x32 = x.detach()            # made contiguous only if a backward pass asks for it
            elif rc != C.GFC_ERR_UNSUPPORTED:
                C.check(rc, "gfc_filter_fwd_pos_nm")
        done = x32 is not None
        stats = None
        mask = None
        if not done:
            x32 = x.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            st = _stream()
            if done:
                pass
            elif src.kind == _SRC_DENSE:
                nb = C.lib.gfc_filter_workspace_bytes(B, N, G, F_, K, E, 0)
                ws = _workspace(nb, dev)
                pflag = prec | src.flags()
                mask = _LSIGF._offer_mask(x, weight, act, B, N, G, F_, K, dev)
                C.check(C.lib.gfc_filter_fwd(C.ptr(x32), C.ptr(src.S), C.ptr(w32), C.ptr(b32), C.ptr(y),
                                             B, N, G, F_, K, E, act, slope, pflag, C.ptr(ws), nb, st),
                        "gfc_filter_fwd")
                mask = mask if (mask is not None and C.lib.gfc_mask_filled()) else None
            elif src.kind == _SRC_POS:
                nb = C.lib.gfc_filter_workspace_bytes(B, N, G, F_, K, 1, 0)
                ws = _workspace(nb, dev)
                if (x.requires_grad or weight.requires_grad) and G in (64, 128) and F_ in (64, 128) and N <= 128:
                    # operand statistics for the backward call of this batch (max |x|: saves its extra pass over x)
                    stats = torch.empty(4, dtype=torch.float32, device=dev)
                    C.lib.gfc_use_stats(C.ptr(stats))
                mask = _LSIGF._offer_mask(x, weight, act, B, N, G, F_, K, dev)
                C.check(C.lib.gfc_filter_fwd_pos(C.ptr(x32), C.ptr(src.pos), src.radius, src.mode,
                                                 C.ptr(w32), C.ptr(b32), C.ptr(y), B, N, G, F_, K,
                                                 act, slope, prec, C.ptr(ws), nb, st), "gfc_filter_fwd_pos")
                mask = mask if (mask is not None and C.lib.gfc_mask_filled()) else None
            else:
                csr = src.csr
                nb = C.lib.gfc_filter_csr_workspace_bytes(B, N, G, F_, K, 0)
                ws = _workspace(nb, dev)
                C.check(C.lib.gfc_filter_csr_fwd(C.ptr(x32), C.ptr(csr.rowptr), C.ptr(csr.colidx),
                                                 C.ptr(csr.vals), csr.nnz_stride, C.ptr(w32), C.ptr(b32),
                                                 C.ptr(y), B, N, G, F_, K, act, slope, prec,
                                                 C.ptr(ws), nb, st), "gfc_filter_csr_fwd")
        ctx.src, ctx.act, ctx.slope, ctx.prec = src, act, slope, prec
        ctx.has_bias = bias is not None
        ctx.stats = stats
        ctx.mask = mask      # signs of y written by the tcgen05 forward kernel: the backward reads them instead of y
        ctx.in_dtypes = (x.dtype, weight.dtype, bias.dtype if bias is not None else None)
        ctx.save_for_backward(x32, w32, y if act != C.ACT_NONE else None)
        return y

    @staticmethod
    def _offer_mask(x, weight, act, B, N, G, F_, K, dev):
        """activation-mask buffer for the forward call that follows (gfc_use_mask), when a backward pass may come"""
        if act == C.ACT_NONE or not (x.requires_grad or weight.requires_grad):
            return None
        nbm = C.lib.gfc_filter_mask_bytes(B, N, G, F_, K)
        if nbm == 0:
            return None
        mask = torch.empty(nbm // 4, dtype=torch.int32, device=dev)
        C.check(C.lib.gfc_use_mask(C.ptr(mask), nbm), "gfc_use_mask")
        return mask

    @staticmethod
    def backward(ctx, dY):
        x32, w32, yout = ctx.saved_tensors
        x32 = x32.contiguous()
        src, act, slope, prec = ctx.src, ctx.act, ctx.slope, ctx.prec
        F_, E, K, G = w32.shape
        B, _, N = x32.shape
        dev = x32.device
        dY = dY.to(torch.float32).contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dX = torch.empty_like(x32) if need_x else None
        dH = torch.empty_like(w32) if need_w else None
        db = torch.empty((F_,), dtype=torch.float32, device=dev) if need_b else None
        with torch.cuda.device(dev):
            st = _stream()
            if src.kind == _SRC_DENSE:
                nb = C.lib.gfc_filter_workspace_bytes(B, N, G, F_, K, E, 1)
                ws = _workspace(nb, dev)
                pflag = prec | src.flags()
                if ctx.mask is not None:
                    C.check(C.lib.gfc_use_mask(C.ptr(ctx.mask), ctx.mask.numel() * 4), "gfc_use_mask")
                C.check(C.lib.gfc_filter_bwd(C.ptr(x32), C.ptr(src.S), C.ptr(w32), C.ptr(yout), C.ptr(dY),
                                             C.ptr(dX), C.ptr(dH), C.ptr(db), B, N, G, F_, K, E,
                                             act, slope, pflag, C.ptr(ws), nb, st), "gfc_filter_bwd")
            elif src.kind == _SRC_POS:
                nb = C.lib.gfc_filter_workspace_bytes(B, N, G, F_, K, 1, 1)
                ws = _workspace(nb, dev)
                if ctx.stats is not None:
                    C.lib.gfc_use_stats(C.ptr(ctx.stats))
                if ctx.mask is not None:
                    C.check(C.lib.gfc_use_mask(C.ptr(ctx.mask), ctx.mask.numel() * 4), "gfc_use_mask")
                C.check(C.lib.gfc_filter_bwd_pos(C.ptr(x32), C.ptr(src.pos), src.radius, src.mode,
                                                 C.ptr(w32), C.ptr(yout), C.ptr(dY), C.ptr(dX), C.ptr(dH),
                                                 C.ptr(db), B, N, G, F_, K, act, slope, prec,
                                                 C.ptr(ws), nb, st), "gfc_filter_bwd_pos")
            else:
                csr = src.csr
                nb = C.lib.gfc_filter_csr_workspace_bytes(B, N, G, F_, K, 1)
                ws = _workspace(nb, dev)
                C.check(C.lib.gfc_filter_csr_bwd(C.ptr(x32), C.ptr(csr.rowptr), C.ptr(csr.colidx), C.ptr(csr.vals),
                                                 C.ptr(csr.rowptr), C.ptr(csr.colidx), C.ptr(csr.vals),
                                                 csr.nnz_stride, C.ptr(w32), C.ptr(yout), C.ptr(dY),
                                                 C.ptr(dX), C.ptr(dH), C.ptr(db), B, N, G, F_, K,
                                                 act, slope, prec, C.ptr(ws), nb, st), "gfc_filter_csr_bwd")
        xd, wd, bd = ctx.in_dtypes
        gx = dX.to(xd) if dX is not None else None
        gw = dH.to(wd) if dH is not None else None
        gb = db.view(F_, 1).to(bd) if db is not None else None
        return gx, gw, gb, None, None, None, None


def graph_filter(x, weight, bias, src, activation=None, negative_slope=0.01, precision="fp32"):
    """functional form; returns the node-major memory ``[B,N,F]``."""
    return _LSIGF.apply(x, weight, bias, src, C.ACTIVATIONS[activation], float(negative_slope),
                        C.PRECISIONS[precision])


class GraphFilterBatch(nn.Module):
    """``GraphFilterBatch(G, F, K, E=1, bias=True)`` — graphML.py:2419.

    x [B,G,Nin] -> y [B,F,Nin] with one GSO per batch element (``addGSO``)."""

    def __init__(self, G, F, K, E=1, bias=True, activation=None, negative_slope=0.01,
                 precision="fp32", reference_dtype=False):
        super().__init__()
        self.G, self.F, self.K, self.E = G, F, K, E
        self.S = None  # no GSO assigned yet (graphML.py:2431)
        self.N = None
        self._src = None
        assert activation in C.ACTIVATIONS, "activation must be one of %r" % list(C.ACTIVATIONS)
        assert precision in C.PRECISIONS, "precision must be one of %r" % list(C.PRECISIONS)
        # the backward recovers the activation mask from the sign of the forward output: needs slope >= 0
        assert negative_slope >= 0, "negative_slope must be >= 0"
        self.activation, self.negative_slope = activation, negative_slope
        self.precision, self.reference_dtype = precision, reference_dtype
        self.weight = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        # U(-s, s), s = 1/sqrt(G*K)  (graphML.py:2442-2447)
        stdv = 1. / math.sqrt(self.G * self.K)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    # ---- graph sources ---------------------------------------------------------
    def addGSO(self, S):
        # same checks, same order, as graphML.py:2449-2456
        assert len(S.shape) == 4
        assert S.shape[1] == self.E
        self.N = S.shape[2]
        assert S.shape[3] == self.N
        self.S = S
        self._src = None  # device copy made lazily in forward

    def addPositions(self, pos, radius, mode="binary_le"):
        """GSO = f(robot positions [B,N,2], communication radius); E must be 1."""
        assert self.E == 1, "a position-built GSO has one edge feature"
        assert len(pos.shape) == 3 and pos.shape[2] == 2
        assert mode in C.GSO_MODES
        _require_cuda(pos, "positions")
        self.N = pos.shape[1]
        self.S = pos  # "GSO stored" for extra_repr
        self._src = _Src(_SRC_POS, pos=pos.detach().to(torch.float32).contiguous(),
                         radius=float(radius), mode=C.GSO_MODES[mode])

    def addSparseGSO(self, csr_or_pos, radius=None, mode="binary_le", max_degree=None):
        """CSR path (kernel (d)) for large sparse swarms; accepts a ``SparseGSO``
        or positions + radius (``max_degree``: sync-free capacity build, see ``build_csr``)."""
        assert self.E == 1
        csr = csr_or_pos if isinstance(csr_or_pos, SparseGSO) else build_csr(csr_or_pos, radius, mode, max_degree)
        self.N = csr.N
        self.S = csr
        self._src = _Src(_SRC_CSR, csr=csr)

    def _source(self, device):
        if self._src is None:
            S = self.S
            assert S is not None, "call addGSO / addPositions before forward"
            _require_cuda(S, "the GSO")
            S32 = S.detach().to(device=device, dtype=torch.float32).contiguous()
            self._src = _Src(_SRC_DENSE, S=S32, binary=_is_binary(S32, self.E, self.G, self.F))
        return self._src

    # ---- forward ----------------------------------------------------------------
    def forward(self, x):
        _require_cuda(x, "x")
        B, G, Nin = x.shape
        assert G == self.G
        N = self.N
        assert Nin <= N
        if Nin < N:  # zero-pad the missing nodes (graphML.py:2464-2468)
            x = torch.cat((x, torch.zeros(B, G, N - Nin, dtype=x.dtype, device=x.device)), dim=2)
        src = self._source(x.device)
        nb = src.S.shape[0] if src.kind == _SRC_DENSE else (src.pos.shape[0] if src.kind == _SRC_POS else src.csr.B)
        assert nb == B, "GSO batch (%d) != x batch (%d)" % (nb, B)
        ymem = graph_filter(x, self.weight, self.bias, src, self.activation, self.negative_slope, self.precision)
        u = ymem.permute(0, 2, 1)  # [B,F,N] view over [B,N,F] memory, as graphML.py:2362
        if Nin < N:
            u = torch.index_select(u, 2, torch.arange(Nin, device=u.device))  # graphML.py:2475-2476
        if self.reference_dtype:
            u = u.double()
        return u

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, " % (
            self.G, self.F) + "filter_taps=%d, " % (
            self.K) + "edge_features=%d, " % (self.E) + \
            "bias=%s, " % (self.bias is not None)
        if self.S is not None:
            reprString += "GSO stored"
        else:
            reprString += "no GSO stored"
        return reprString


class GraphFilter(GraphFilterBatch):
    """``GraphFilter(G, F, K, E=1, bias=True)`` — graphML.py:1111 (``LSIGF``, graphML.py:48).

    One GSO ``S [E,N,N]`` shared by every batch element: x [B,G,Nin] -> y [B,F,Nin]
    in x's dtype (``LSIGF`` has none of ``BatchLSIGF``'s float64 casts). The shared
    GSO is presented to the batch kernels as ``B`` identical graphs."""

    def __init__(self, G, F, K, E=1, bias=True, activation=None, negative_slope=0.01,
                 precision="fp32"):
        super().__init__(G, F, K, E, bias, activation, negative_slope, precision, reference_dtype=False)
        self._srcB = 0

    def addGSO(self, S):
        # same checks, same order, as graphML.py:1190-1197
        assert len(S.shape) == 3
        assert S.shape[0] == self.E
        self.N = S.shape[1]
        assert S.shape[2] == self.N
        self.S = S
        self._src, self._srcB = None, 0

    def addPositions(self, pos, radius, mode="binary_le"):
        raise TypeError("GraphFilter holds one GSO [E,N,N]; use GraphFilterBatch for per-sample positions")

    addSparseGSO = addPositions

    def _source(self, device, B=None):
        if self._src is None or self._srcB != B:
            S = self.S
            assert S is not None, "call addGSO before forward"
            _require_cuda(S, "the GSO")
            S32 = S.detach().to(device=device, dtype=torch.float32)
            # ONE copy of S [E,N,N]; the kernels index it with a zero batch stride (GFC_PREC_FLAG_SHARED_GSO)
            self._src = _Src(_SRC_DENSE, S=S32.contiguous(), binary=_is_binary(S32, self.E, self.G, self.F), shared=True)
            self._srcB = B
        return self._src

    def forward(self, x):
        _require_cuda(x, "x")
        B, G, Nin = x.shape
        assert G == self.G
        N = self.N
        assert Nin <= N
        if Nin < N:  # graphML.py:1205-1209
            x = torch.cat((x, torch.zeros(B, G, N - Nin, dtype=x.dtype, device=x.device)), dim=2)
        src = self._source(x.device, B)
        ymem = graph_filter(x, self.weight, self.bias, src, self.activation, self.negative_slope, self.precision)
        u = ymem.permute(0, 2, 1)
        if Nin < N:
            u = torch.index_select(u, 2, torch.arange(Nin, device=u.device))  # graphML.py:1216-1217
        return u.to(x.dtype)


class GraphFilterBatchGSO(GraphFilterBatch):
    """``GraphFilterBatchGSO(G, F, K, E=1, bias=True)`` — graphML.py:2174 (``matrixPowersBatch`` :2063,
    ``batchLSIGF`` :2107): one GSO per batch element given as ``[B,N,N]`` or ``[B,E,N,N]``.

    The reference precomputes the powers ``S_b^k`` and contracts ``x_b S_b^k`` with the taps; that is the filter of
    ``GraphFilterBatch`` evaluated in another order, so the same fused kernels serve it (the powers are never
    formed on the device path).  Like the reference it runs in x's dtype (no float64 casts), requires
    ``x.shape == (B, G, N)`` exactly (no zero-padding) and, given a GSO of another shape, keeps the previous one."""

    def __init__(self, G, F, K, E=1, bias=True, activation=None, negative_slope=0.01, precision="fp32"):
        super().__init__(G, F, K, E, bias, activation, negative_slope, precision, reference_dtype=False)
        self.B = None

    def addGSO(self, S):
        # graphML.py:2229-2244: 3-d -> one edge feature; 4-d must carry E edge features; anything else is ignored
        if len(S.shape) == 3 and S.shape[1] == S.shape[2]:
            self.S = S.unsqueeze(1)
        elif len(S.shape) == 4 and S.shape[1] == self.E and S.shape[2] == S.shape[3]:
            self.S = S
        self.N = self.S.shape[2]
        self.B = self.S.shape[0]
        self._src = None

    @property
    def SK(self):
        """the reference's attribute ``SK [B,E,K,N,N]`` (graphML.py:2246), formed on demand with torch ops"""
        S = self.S
        cur = torch.eye(self.N, dtype=S.dtype, device=S.device).repeat(self.B, self.E, 1, 1)
        out = [cur]
        for _ in range(1, self.K):
            cur = torch.matmul(cur, S)
            out.append(cur)
        return torch.stack(out, dim=2)

    def addPositions(self, pos, radius, mode="binary_le"):
        super().addPositions(pos, radius, mode)
        self.B = pos.shape[0]

    def forward(self, x):
        # batchLSIGF's asserts (graphML.py:2149-2154)
        assert x.shape[0] == self.B
        assert x.shape[1] == self.G
        assert x.shape[2] == self.N
        _require_cuda(x, "x")
        src = self._source(x.device)
        ymem = graph_filter(x, self.weight, self.bias, src, self.activation, self.negative_slope, self.precision)
        return ymem.permute(0, 2, 1).to(x.dtype)

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, " % (
            self.G, self.F) + "filter_taps=%d, " % (
            self.K) + "edge_features=%d, " % (self.E) + \
            "bias=%s, " % (self.bias is not None)
        if self.S is not None:
            reprString += "GSO stored: number_nodes=%d, batch_size=%d" % (self.N, self.B)
        else:
            reprString += "no GSO stored"
        return reprString
