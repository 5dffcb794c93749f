"""Recurrent graph-filter layers of the reference as compositions of the fused filter kernels.

Reference classes (utils/graphUtils/graphML.py): ``GraphFilterRNNBatch`` :2491-2654, ``torchpermul`` :2656-2679,
``GraphFilterMoRNNBatch`` :2681-2835, ``GraphFilterL2ShareBatch`` :2837-2987.  Same constructors
``(G, H, F, K, E=1, bias=True)``, same parameter names / shapes (``weight_A [H,E,K,G]``, ``weight_B``, ``weight_D``,
``bias_A/B/D``), same initialisation laws, same ``addGSO`` asserts, ``updateHiddenState`` / ``detachHiddenState``,
same zero-pad / slice semantics, same ``extra_repr``.

How they map onto kernel (b)/(c):

* ``GraphFilterRNNBatch``: ``h' = ReLU(LSIGF(A, S, x) + LSIGF(B, S, h))`` is ONE filter over the channel
  concatenation ``[x ; h]`` with taps ``[A | B]`` (concatenated along the input-feature axis) and bias
  ``b_A + b_B`` — one fused launch with the ReLU in the epilogue instead of two filters, an add and an activation
  pass; ``u = LSIGF(D, S, h')`` consumes the node-major ``h'`` in place.  Autograd flows through ``torch.cat``.
* ``GraphFilterMoRNNBatch`` / ``GraphFilterL2ShareBatch``: only the input branch is a graph filter; the hidden
  and output branches are the reference's ``torchpermul`` — an elementwise broadcast product (``torch.mul``, not
  a matmul; it only broadcasts when ``N == H`` etc.), kept as the same torch expression.

One deviation, on purpose: with ``bias=False`` the reference constructor raises ``AttributeError`` (it registers
``'bias'`` and then touches ``bias_A``, graphML.py:2560-2570); here the three biases are registered as ``None``.
"""
import math

import torch
import torch.nn as nn

from . import _cabi as C
from .graph_filter import GraphFilterBatch, graph_filter, _require_cuda, _SRC_DENSE, _SRC_POS


def torchpermul(h, x, b=None):
    """graphML.py:2656-2679 — ``(x.permute(0,2,1) * h.permute(1,0)).permute(0,2,1) (+ b)``, an elementwise product."""
    y = torch.mul(x.permute(0, 2, 1), h.permute(1, 0)).permute(0, 2, 1)
    if b is not None:
        y = y + b
    return y


class _RecurrentBase(nn.Module):
    """shared plumbing: GSO sources (same surface as GraphFilterBatch), hidden-state handling, repr"""

    def __init__(self, G, H, F, K, E, precision, reference_dtype):
        super().__init__()
        self.G, self.F, self.H, self.K, self.E = G, F, H, K, E
        self.S = None
        self.N = None
        assert precision in C.PRECISIONS
        self.precision, self.reference_dtype = precision, reference_dtype
        self._src = None

    # ---- graph sources: GraphFilterBatch's own methods (graphML.py:2585-2592 asserts = :2449-2456; plus the additive
    # position / CSR builders), reused as plain functions so the classes cannot drift from the base layer ----------
    addGSO = GraphFilterBatch.addGSO
    addPositions = GraphFilterBatch.addPositions
    addSparseGSO = GraphFilterBatch.addSparseGSO
    _source = GraphFilterBatch._source

    def updateHiddenState(self, hiddenState):
        self.hiddenState = hiddenState

    def detachHiddenState(self):
        self.hiddenState.detach_()
        self.hiddenStateNext.detach_()

    def _pad(self, x):
        B, Fx, Nin = x.shape
        if Nin < self.N:
            x = torch.cat((x, torch.zeros(B, Fx, self.N - Nin, dtype=x.dtype, device=x.device)), dim=2)
        return x, Nin

    def _filter(self, x, weight, bias, activation=None):
        """[B,*,N] -> [B,F,N] view over node-major memory (float64 with reference_dtype, as BatchLSIGF returns)"""
        src = self._source(x.device)
        nb = src.S.shape[0] if src.kind == _SRC_DENSE else (src.pos.shape[0] if src.kind == _SRC_POS else src.csr.B)
        assert nb == x.shape[0], "GSO batch (%d) != x batch (%d)" % (nb, x.shape[0])
        u = graph_filter(x, weight, bias, src, activation, 0.0, self.precision).permute(0, 2, 1)
        return u.double() if self.reference_dtype else u

    def _finish(self, u, Nin):
        if Nin < self.N:
            u = torch.index_select(u, 2, torch.arange(Nin, device=u.device))
        return u

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, hidden_features=%d, " % (
            self.G, self.F, self.H) + "filter_taps=%d, " % (
            self.K) + "edge_features=%d, " % (self.E) + \
            "bias=%s, " % (self.bias_D is not None)
        if self.S is not None:
            reprString += "GSO stored"
        else:
            reprString += "no GSO stored"
        return reprString

    def _uniform(self, w, b, fan):
        stdv = 1. / math.sqrt(fan)
        w.data.uniform_(-stdv, stdv)
        if b is not None:
            b.data.uniform_(-stdv, stdv)


class GraphFilterRNNBatch(_RecurrentBase):
    """``GraphFilterRNNBatch(G, H, F, K, E=1, bias=True)`` — graphML.py:2491.

    ``h' = ReLU(LSIGF(A,S,x) + LSIGF(B,S,h))``; ``u = LSIGF(D,S,h')``; x [B,G,Nin], h [B,H,N] -> u [B,F,Nin]."""

    def __init__(self, G, H, F, K, E=1, bias=True, precision="fp32", reference_dtype=False):
        super().__init__(G, H, F, K, E, precision, reference_dtype)
        self.weight_A = nn.parameter.Parameter(torch.Tensor(H, E, K, G))
        self.weight_B = nn.parameter.Parameter(torch.Tensor(H, E, K, H))
        self.weight_D = nn.parameter.Parameter(torch.Tensor(F, E, K, H))
        for name, n in (("bias_A", H), ("bias_B", H), ("bias_D", F)):
            if bias:
                setattr(self, name, nn.parameter.Parameter(torch.Tensor(n, 1)))
            else:
                self.register_parameter(name, None)
        self.reset_parameters()

    def reset_parameters(self):
        # graphML.py:2572-2588
        self._uniform(self.weight_A, self.bias_A, self.G * self.K)
        self._uniform(self.weight_B, self.bias_B, self.H * self.K)
        self._uniform(self.weight_D, self.bias_D, self.H * self.K)

    def forward(self, x):
        _require_cuda(x, "x")
        assert x.shape[1] == self.G
        x, Nin = self._pad(x)
        h = self.hiddenState
        assert h.shape[0] == x.shape[0] and h.shape[1] == self.H and h.shape[2] == self.N
        # u_a + u_b as one filter over [x ; h] with taps [A | B] (flatten order (e,k,g) keeps g innermost)
        w_ab = torch.cat((self.weight_A, self.weight_B), dim=3)
        b_ab = (self.bias_A + self.bias_B) if self.bias_A is not None else None
        xh = torch.cat((x.to(torch.float32), h.to(torch.float32)), dim=1)
        self.hiddenStateNext = self._filter(xh, w_ab, b_ab, "relu")
        u = self._filter(self.hiddenStateNext, self.weight_D, self.bias_D)
        self.updateHiddenState(self.hiddenStateNext)
        return self._finish(u, Nin)


class GraphFilterMoRNNBatch(_RecurrentBase):
    """``GraphFilterMoRNNBatch(G, H, F, K, E=1, bias=True)`` — graphML.py:2681: graph filter on the input branch,
    ``torchpermul`` (elementwise) on the hidden and output branches."""

    def __init__(self, G, H, F, K, E=1, bias=True, precision="fp32", reference_dtype=False):
        super().__init__(G, H, F, K, E, precision, reference_dtype)
        self.weight_A = nn.parameter.Parameter(torch.Tensor(H, E, K, G))
        self.weight_B = nn.parameter.Parameter(torch.Tensor(H, H))
        self.weight_D = nn.parameter.Parameter(torch.Tensor(F, H))
        for name, n in (("bias_A", H), ("bias_B", H), ("bias_D", F)):
            if bias:
                setattr(self, name, nn.parameter.Parameter(torch.Tensor(n, 1)))
            else:
                self.register_parameter(name, None)
        self.reset_parameters()

    def reset_parameters(self):
        # graphML.py:2751-2765
        self._uniform(self.weight_A, self.bias_A, self.G * self.K)
        self._uniform(self.weight_B, self.bias_B, self.H)
        self._uniform(self.weight_D, self.bias_D, self.H)

    def forward(self, x):
        _require_cuda(x, "x")
        assert x.shape[1] == self.G
        x, Nin = self._pad(x)
        u_a = self._filter(x, self.weight_A, self.bias_A)
        u_b = torchpermul(self.weight_B, self.hiddenState, self.bias_B)
        self.hiddenStateNext = torch.relu(u_a + u_b)
        u = torchpermul(self.weight_D, self.hiddenStateNext, self.bias_D)
        self.updateHiddenState(self.hiddenStateNext)
        return self._finish(u, Nin)


class GraphFilterL2ShareBatch(GraphFilterMoRNNBatch):
    """``GraphFilterL2ShareBatch(G, H, F, K, E=1, bias=True)`` — graphML.py:2837; the reference class is a
    statement-for-statement twin of ``GraphFilterMoRNNBatch`` (same parameters, same forward)."""
