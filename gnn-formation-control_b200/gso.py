"""Position -> graph shift operator on the GPU (kernel (a) + the CSR builder).

Host-side mirror of the reference builders:
  * ``Scene.readADjMatrix(MaxRange)``                       scene.py:140-154  -> mode "binary_le"
  * ``multiRobotSim.computeAdjacencyMatrix_fixedCommRadius`` utils/multirobotsim_dcenlocal.py:291-317
                                                             -> mode "sym_norm_lt"
The adjacency / 1-hop mask is bit-identical to the reference's float64 result.
"""
import torch

from . import _cabi as C


def _stream():
    return C.ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_pos(pos):
    assert pos.dim() == 3 and pos.shape[2] == 2, "positions must be [B, N, 2]"
    if not pos.is_cuda:
        raise RuntimeError("gnnfc: positions must live on a CUDA device (no CPU fallback)")
    return pos.detach().to(torch.float32).contiguous()


def build_gso(pos, radius, mode="binary_le", want_adj=True, want_S=True):
    """pos [B,N,2] (cuda) -> (adj uint8 [B,N,N] | None, S float32 [B,N,N] | None)."""
    pos = _as_pos(pos)
    B, N, _ = pos.shape
    adj = torch.empty((B, N, N), dtype=torch.uint8, device=pos.device) if want_adj else None
    S = torch.empty((B, N, N), dtype=torch.float32, device=pos.device) if want_S else None
    with torch.cuda.device(pos.device):
        C.check(C.lib.gfc_gso_build(C.ptr(pos), B, N, float(radius), C.GSO_MODES[mode],
                                    C.ptr(adj), C.ptr(S), _stream()), "gfc_gso_build")
    return adj, S


class SparseGSO:
    """Batched CSR gather lists of a position-built (symmetric) GSO.

    rowptr int32 [B, N+1] (per-graph offsets), colidx int32 [B, nnz_stride]
    ascending within a row, vals float32 [B, nnz_stride] or None (all ones).
    """

    def __init__(self, rowptr, colidx, vals, nnz_stride, N, overflow=None):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.nnz_stride, self.N = int(nnz_stride), int(N)
        self.B = rowptr.shape[0]
        self.overflow = overflow     # int32 [1] device word of the sync-free builder (None: exact sizing)

    def check(self):
        """SYNCHRONISING: raises if a graph needed more than the ``max_degree`` capacity it was built with"""
        if self.overflow is not None:
            need = int(self.overflow.item())
            if need > self.nnz_stride:
                raise RuntimeError("SparseGSO: a graph has %d edges, capacity is %d (N=%d): rebuild with max_degree >= %d"
                                   % (need, self.nnz_stride, self.N, -(-need // max(self.N, 1))))
        return True

    def to_dense(self):
        B, N = self.B, self.N
        S = torch.zeros((B, N, N), dtype=torch.float32, device=self.rowptr.device)
        rp = self.rowptr.cpu()
        ci = self.colidx.cpu()
        vv = self.vals.cpu() if self.vals is not None else None
        Sc = S.cpu()
        for b in range(B):
            for n in range(N):
                lo, hi = int(rp[b, n]), int(rp[b, n + 1])
                cols = ci[b, lo:hi].long()
                Sc[b, cols, n] = vv[b, lo:hi] if vv is not None else 1.0  # list of n holds S[m, n]
        return Sc.to(self.rowptr.device)


def build_csr(pos, radius, mode="binary_le", max_degree=None):
    """CSR gather lists of the radius graph, built by ``gfc_csr_build`` (one launch, cell list).

    ``max_degree=None``: exact sizing — a rowptr-only pass, ONE host read of the largest per-graph nnz, then the fill.
    ``max_degree=k``: sync-free — per-graph capacity ``N*k``, nothing leaves the device (CUDA-graph capturable); a graph
    with more edges sets ``csr.overflow`` and loses the surplus, ``csr.check()`` raises in that case.
    Graphs too large for the fused builder (N above ~8000) take the three-kernel count / scan / fill path."""
    pos = _as_pos(pos)
    B, N, _ = pos.shape
    dev = pos.device
    m = C.GSO_MODES[mode]
    norm = m == C.GSO_SYM_NORM_LT
    rowptr = torch.empty((B, N + 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _stream()
        if max_degree is not None:
            nnz_stride = max(4, (N * min(int(max_degree), max(N - 1, 1)) + 3) // 4 * 4)
            colidx = torch.empty((B, nnz_stride), dtype=torch.int32, device=dev)
            vals = torch.empty((B, nnz_stride), dtype=torch.float32, device=dev) if norm else None
            overflow = torch.zeros(1, dtype=torch.int32, device=dev)
            rc = C.lib.gfc_csr_build(C.ptr(pos), B, N, float(radius), m, C.ptr(rowptr), nnz_stride, C.ptr(colidx),
                                     C.ptr(vals), C.ptr(overflow), st)
            if rc == C.GFC_OK:
                return SparseGSO(rowptr, colidx, vals, nnz_stride, N, overflow)
            if rc != C.GFC_ERR_UNSUPPORTED:
                C.check(rc, "gfc_csr_build")
        rc = C.lib.gfc_csr_build(C.ptr(pos), B, N, float(radius), m, C.ptr(rowptr), 0, None, None, None, st)
        fused = rc == C.GFC_OK
        if not fused:
            if rc != C.GFC_ERR_UNSUPPORTED:
                C.check(rc, "gfc_csr_build")
            deg = torch.empty((B, N), dtype=torch.int32, device=dev)
            C.check(C.lib.gfc_csr_count(C.ptr(pos), B, N, float(radius), m, C.ptr(deg), st), "gfc_csr_count")
            C.check(C.lib.gfc_csr_scan(C.ptr(deg), B, N, C.ptr(rowptr), st), "gfc_csr_scan")
        nnz_stride = max(4, int(rowptr[:, N].max().item())) if B > 0 else 4
        nnz_stride = (nnz_stride + 3) // 4 * 4
        colidx = torch.zeros((B, nnz_stride), dtype=torch.int32, device=dev)
        vals = torch.zeros((B, nnz_stride), dtype=torch.float32, device=dev) if norm else None
        if fused:
            C.check(C.lib.gfc_csr_build(C.ptr(pos), B, N, float(radius), m, C.ptr(rowptr), nnz_stride, C.ptr(colidx),
                                        C.ptr(vals), None, st), "gfc_csr_build")
        else:
            C.check(C.lib.gfc_csr_fill(C.ptr(pos), B, N, float(radius), m, C.ptr(rowptr), nnz_stride,
                                       C.ptr(colidx), C.ptr(vals), st), "gfc_csr_fill")
    return SparseGSO(rowptr, colidx, vals, nnz_stride, N)


def build_csr_three_pass(pos, radius, mode="binary_le"):
    """the count -> scan -> fill builder (O(N^2) pair walk, any N); kept for graphs beyond the fused builder and as
    the cross-check of ``gfc_csr_build`` in the tests"""
    pos = _as_pos(pos)
    B, N, _ = pos.shape
    dev = pos.device
    m = C.GSO_MODES[mode]
    deg = torch.empty((B, N), dtype=torch.int32, device=dev)
    rowptr = torch.empty((B, N + 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _stream()
        C.check(C.lib.gfc_csr_count(C.ptr(pos), B, N, float(radius), m, C.ptr(deg), st), "gfc_csr_count")
        C.check(C.lib.gfc_csr_scan(C.ptr(deg), B, N, C.ptr(rowptr), st), "gfc_csr_scan")
        nnz_stride = max(4, int(rowptr[:, N].max().item())) if B > 0 else 4
        nnz_stride = (nnz_stride + 3) // 4 * 4
        colidx = torch.zeros((B, nnz_stride), dtype=torch.int32, device=dev)
        vals = torch.zeros((B, nnz_stride), dtype=torch.float32, device=dev) if m == C.GSO_SYM_NORM_LT else None
        C.check(C.lib.gfc_csr_fill(C.ptr(pos), B, N, float(radius), m, C.ptr(rowptr), nnz_stride,
                                   C.ptr(colidx), C.ptr(vals), st), "gfc_csr_fill")
    return SparseGSO(rowptr, colidx, vals, nnz_stride, N)
