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

    def __init__(self, rowptr, colidx, vals, nnz_stride, N):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.nnz_stride, self.N = int(nnz_stride), int(N)
        self.B = rowptr.shape[0]

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


def build_csr(pos, radius, mode="binary_le"):
    """Two-pass CSR build (count -> scan -> fill); one host sync to size colidx."""
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
