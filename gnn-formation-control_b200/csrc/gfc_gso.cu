// gfc_gso.cu — kernel (a): robot positions -> graph shift operator, plus the CSR
// builder used by the large-swarm path.
//
// Reference behaviour reproduced (bit-exact mask):
//   Scene.readADjMatrix                      scene.py:140-154   (d <= R, zero diag, {0,1})
//   computeAdjacencyMatrix_fixedCommRadius   utils/multirobotsim_dcenlocal.py:291-317
//                                            (d < R, zero diag, D^-1/2 W D^-1/2, deg==0 guard)
// The reference evaluates the distance in float64; so does this kernel (6 fp64
// ops per pair), comparing the squared distance with a host-computed threshold
// that is equivalent to comparing the correctly rounded sqrt with R.
#include "gfc_common.cuh"
#include <math.h>

namespace gfc {

double squared_threshold(double R, bool inclusive) {
  if (isnan(R)) return -1.0;
  if (inclusive) {
    if (R < 0) return -1.0;
    if (isinf(R)) return INFINITY;
    double s = R * R;
    if (isinf(s)) return 1.7976931348623157e308;
    while (sqrt(s) <= R) s = nextafter(s, INFINITY);
    while (s >= 0 && sqrt(s) > R) s = nextafter(s, -INFINITY);
    return s;
  }
  if (R <= 0) return -1.0;
  if (isinf(R)) return 1.7976931348623157e308;
  double s = R * R;
  if (isinf(s)) return 1.7976931348623157e308;
  while (sqrt(s) < R) s = nextafter(s, INFINITY);
  while (s > 0 && !(sqrt(s) < R)) s = nextafter(s, -INFINITY);
  if (!(sqrt(s) < R)) return -1.0;
  return s;
}

// One CTA handles `gpc` consecutive graphs.  smem: [isd: gpc*N doubles (NORM only)]
// [positions: gpc*N*2 floats].
template <bool NORM>
__global__ void __launch_bounds__(256)
gso_build_kernel(const float* __restrict__ pos, int B, int N, double thr, int gpc,
                 uint8_t* __restrict__ adj, float* __restrict__ S) {
  extern __shared__ double smem_d[];
  double* isd = smem_d;
  float* sp = reinterpret_cast<float*>(smem_d + (NORM ? (size_t)gpc * N : 0));
  const int tid = threadIdx.x, nt = blockDim.x;
  const int b0 = blockIdx.x * gpc;
  const int gcount = min(gpc, B - b0);
  const int nn = gcount * N;
  const float* gp = pos + (size_t)b0 * N * 2;
  for (int i = tid; i < nn * 2; i += nt) sp[i] = gp[i];
  __syncthreads();
  if (NORM) {
    for (int r = tid; r < nn; r += nt) {
      const int j = r / N, i = r - j * N;
      const float xi = sp[2 * r], yi = sp[2 * r + 1];
      int deg = 0;
      for (int m = 0; m < N; ++m) {
        const int q = j * N + m;
        deg += (m != i) && (sqdist64(xi, yi, sp[2 * q], sp[2 * q + 1]) <= thr);
      }
      isd[r] = inv_sqrt_deg(deg);
    }
    __syncthreads();
  }
  const size_t base = (size_t)b0 * N * N;
  const int total = nn * N;
  for (int o = tid; o < total; o += nt) {
    const int r = o / N, n2 = o - r * N;
    const int j = r / N, i = r - j * N;
    const int q = j * N + n2;
    const bool a = (i != n2) && (sqdist64(sp[2 * r], sp[2 * r + 1], sp[2 * q], sp[2 * q + 1]) <= thr);
    if (adj) adj[base + o] = a ? 1 : 0;
    if (S) {
      float v = a ? 1.f : 0.f;
      if (NORM) v = a ? (float)__dmul_rn(isd[r], isd[q]) : 0.f;
      S[base + o] = v;
    }
  }
}

// CSR build, pass 1 and 3 share one loop: a CTA owns 256 rows of ONE graph and walks over all nodes of that graph
// staged in shared memory (128-bit broadcasts), deciding each pair with the fp32 screen of the fused kernels
// (s < thr_lo: inside, s > thr_hi: outside) and the exact fp64 rule of (a) only inside the rounding band — the same
// bit-exact decisions as sqdist64(...) <= thr everywhere, at a third of the instruction count.
static __device__ __noinline__ bool csr_pair_exact(float xi, float yi, float xj, float yj, double thr) {
  return sqdist64(xi, yi, xj, yj) <= thr;
}
constexpr int kCsrChunk = 2048;   // nodes staged per pass (16 KB)

template <bool FILL, bool NORM>
__global__ void __launch_bounds__(256)
csr_rows_kernel(const float* __restrict__ pos, int N, double thr, float thr_lo, float thr_hi,
                int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr, long long nnz_stride,
                int32_t* __restrict__ colidx, float* __restrict__ vals) {
  __shared__ float2 sp[kCsrChunk];
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const bool valid = i < N;
  const float2* gp = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
  const float2 pi = valid ? __ldg(gp + i) : make_float2(0.f, 0.f);
  const int32_t* rp = FILL ? rowptr + (size_t)b * (N + 1) : nullptr;
  int d = 0;
  long long w = 0;
  double isd_i = 0.0;
  if (FILL && valid) {
    w = (long long)b * nnz_stride + rp[i];
    if (NORM) isd_i = inv_sqrt_deg(rp[i + 1] - rp[i]);
  }
  for (int m0 = 0; m0 < N; m0 += kCsrChunk) {
    const int cnt = min(kCsrChunk, N - m0);
    __syncthreads();
    for (int q = threadIdx.x; q < cnt; q += 256) sp[q] = __ldg(gp + m0 + q);
    __syncthreads();
    if (!valid) continue;
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) {
      const float2 pj = sp[q];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y;
      const float sq = fmaf(dx, dx, dy * dy);
      bool e = sq < thr_lo;
      if (!e && !(sq > thr_hi)) e = csr_pair_exact(pi.x, pi.y, pj.x, pj.y, thr);   // rounding band: exact rule
      e = e && (m0 + q != i);
      if (FILL) {
        if (e) {
          const int m = m0 + q;
          colidx[w] = m;
          if (vals) vals[w] = NORM ? (float)__dmul_rn(inv_sqrt_deg(rp[m + 1] - rp[m]), isd_i) : 1.f;
          ++w;
        }
      } else {
        d += e ? 1 : 0;
      }
    }
  }
  if (!FILL && valid) deg[(size_t)b * N + i] = d;
}

// rowptr[b, 0..N] = exclusive scan of deg[b, :]; one CTA per graph.
__global__ void __launch_bounds__(256)
csr_scan_kernel(const int32_t* __restrict__ deg, int N, int32_t* __restrict__ rowptr) {
  __shared__ int part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int chunk = (N + 255) / 256;
  const int lo = min(N, tid * chunk), hi = min(N, lo + chunk);
  const int32_t* d = deg + (size_t)b * N;
  int32_t* rp = rowptr + (size_t)b * (N + 1);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += d[i];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { int v = part[i]; part[i] = run; run += v; }
    rp[N] = run;
  }
  __syncthreads();
  int run = part[tid];
  for (int i = lo; i < hi; ++i) { rp[i] = run; run += d[i]; }
}


// ---------------------------------------------------------------------------------------------------------------
// CSR build in ONE launch: one CTA per graph, cell list instead of the O(N^2) pair walk.
//   1. positions -> shared memory, bounding box (block reduction);
//   2. uniform grid with cell edge >= 1.001 R (so every neighbour lies in the 3x3 cells around a node; the margin
//      absorbs the fp32 rounding of the cell index, the index map is monotone), at most 64 x 64 cells;
//   3. counting sort of the nodes by cell: shared-memory atomics give an arbitrary slot, the final position is the
//      node's RANK among the indices of its cell, so every cell list is ascending — deterministic;
//   4. thread = row: count the neighbours among the <= 9 cell lists (fp32 screen, exact fp64 rule inside the rounding
//      band — the bit-exact decision of kernel (a)), block scan -> rowptr;
//   5. thread = row: 9-way merge of the (ascending) cell lists emits the row's columns in ascending order.
// nnz_stride is a CAPACITY: if a graph needs more, *overflow = max(*overflow, needed) and the surplus edges are
// dropped (the caller checks the flag; gnnfc.SparseGSO.check()).  colidx == NULL: rowptr only (sizing pass).
constexpr int kCellDim = 64;
constexpr int kCellMax = kCellDim * kCellDim;

struct CellRanges { int lo[3], hi[3]; };

template <bool NORM>
__global__ void __launch_bounds__(1024)
csr_build_fused_kernel(const float* __restrict__ pos, int N, double thr, float thr_lo, float thr_hi, float cell0,
                       int32_t* __restrict__ rowptr, long long nnz_stride, int32_t* __restrict__ colidx,
                       float* __restrict__ vals, int* __restrict__ overflow, int stage_cap) {
  extern __shared__ unsigned char csr_smem[];
  float2* sp = reinterpret_cast<float2*>(csr_smem);          // [N] positions
  int* cellstart = reinterpret_cast<int*>(sp + N);             // [kCellMax + 1]
  int* cid = cellstart + kCellMax + 1;                         // [N] cell of node
  int* lst = cid + N;                                          // [N] nodes ordered by (cell, index)
  int* aux = lst + N;                                          // [N] slot -> degree
  int* un = aux + N;                                           // [N] unsorted cell lists -> local rowptr
  int* stage = un + N;                                         // [stage_cap] the graph's matches, row by row, unsorted
  __shared__ float redf[4][32];
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const float2* gp = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
  int32_t* rp = rowptr + (size_t)b * (N + 1);

  // 1. load + bounding box
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = tid; i < N; i += nt) {
    const float2 p = __ldg(gp + i);
    sp[i] = p;
    mnx = fminf(mnx, p.x); mxx = fmaxf(mxx, p.x); mny = fminf(mny, p.y); mxy = fmaxf(mxy, p.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) { redf[0][wid] = mnx; redf[1][wid] = mxx; redf[2][wid] = mny; redf[3][wid] = mxy; }
  __syncthreads();
  mnx = INFINITY; mny = INFINITY; mxx = -INFINITY; mxy = -INFINITY;
  for (int w = 0; w < nw; ++w) {
    mnx = fminf(mnx, redf[0][w]); mxx = fmaxf(mxx, redf[1][w]); mny = fminf(mny, redf[2][w]); mxy = fmaxf(mxy, redf[3][w]);
  }
  // 2. grid (all threads compute the same values).  Non-finite extents (inf / nan positions) collapse to one cell
  //    per axis: still correct, every pair is then tested.
  float ex = mxx - mnx, ey = mxy - mny;
  if (!(ex >= 0.f) || !(ex < 3.0e38f)) ex = 0.f;
  if (!(ey >= 0.f) || !(ey < 3.0e38f)) ey = 0.f;
  const bool one_cell = !(cell0 > 0.f) || !(cell0 < 3.0e38f) || !(mnx > -3.0e38f) || !(mny > -3.0e38f) ||
                        (mxx - mnx != ex) || (mxy - mny != ey);
  const float cx = fmaxf(cell0, ex * (1.f / (kCellDim - 0.5f)));
  const float cy = fmaxf(cell0, ey * (1.f / (kCellDim - 0.5f)));
  const int gx = one_cell ? 1 : min(kCellDim, (int)__fdiv_rn(ex, cx) + 1);
  const int gy = one_cell ? 1 : min(kCellDim, (int)__fdiv_rn(ey, cy) + 1);
  const int ncell = gx * gy;
  for (int c = tid; c <= ncell; c += nt) cellstart[c] = 0;
  __syncthreads();
  // 3. counting sort by cell
  for (int i = tid; i < N; i += nt) {
    int c = 0;
    if (!one_cell) {
      const float2 p = sp[i];
      int ix = (int)__fdiv_rn(p.x - mnx, cx), iy = (int)__fdiv_rn(p.y - mny, cy);   // NaN -> 0
      ix = max(0, min(gx - 1, ix)); iy = max(0, min(gy - 1, iy));
      c = iy * gx + ix;
    }
    cid[i] = c;
    aux[i] = atomicAdd(&cellstart[c + 1], 1);      // counts live one slot up: the scan below turns them into starts
  }
  __syncthreads();
  {  // inclusive scan of cellstart[1..ncell] (counts) in place -> cellstart[c] = start of cell c
    int carry = 0;
    for (int base = 1; base <= ncell; base += nt) {
      const int c = base + tid;
      const int v = c <= ncell ? cellstart[c] : 0;
      int sc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
      if (lane == 31) wsum[wid] = sc;
      __syncthreads();
      if (wid == 0) {
        int ws = lane < nw ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
        wsum[lane] = ws;
      }
      __syncthreads();
      const int pre = (wid ? wsum[wid - 1] : 0) + carry;
      if (c <= ncell) cellstart[c] = pre + sc;
      carry += wsum[nw - 1];
      __syncthreads();
    }
  }
  // now cellstart[c] = number of nodes in cells < c ... shifted: cellstart[c+1] = inclusive end of cell c, [0] = 0
  for (int i = tid; i < N; i += nt) un[cellstart[cid[i]] + aux[i]] = i;
  __syncthreads();
  for (int i = tid; i < N; i += nt) {
    const int c = cid[i], s0 = cellstart[c], s1 = cellstart[c + 1];
    int r = 0;
    for (int q = s0; q < s1; ++q) r += un[q] < i;
    lst[s0 + r] = i;
  }
  __syncthreads();

  // 4. degrees
  auto ranges = [&](int c, CellRanges& R) {
    const int iy = c / gx, ix = c - iy * gx;
    const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gx - 1);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int y2 = iy + d - 1;
      const bool ok = y2 >= 0 && y2 < gy;
      R.lo[d] = ok ? cellstart[y2 * gx + x0] : 0;
      R.hi[d] = ok ? cellstart[y2 * gx + x1 + 1] : 0;
    }
  };
  auto edge = [&](float2 pi, int i, int j) -> bool {
    const float2 pj = sp[j];
    const float dx = pi.x - pj.x, dy = pi.y - pj.y;
    const float sq = fmaf(dx, dx, dy * dy);
    bool e = sq < thr_lo;
    if (!e && !(sq > thr_hi)) e = csr_pair_exact(pi.x, pi.y, pj.x, pj.y, thr);
    return e && j != i;
  };
  for (int i = tid; i < N; i += nt) {
    CellRanges R;
    ranges(cid[i], R);
    const float2 pi = sp[i];
    int d = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      for (int q = R.lo[k]; q < R.hi[k]; ++q) d += edge(pi, i, lst[q]) ? 1 : 0;
    aux[i] = d;
  }
  __syncthreads();
  // rowptr = exclusive scan of aux (kept in un[] for the fill pass)
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < N; base += nt) {
    const int i = base + tid;
    const int v = i < N ? aux[i] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
    if (lane == 31) wsum[wid] = sc;
    __syncthreads();
    if (wid == 0) {
      int ws = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
      wsum[lane] = ws;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = (wid ? wsum[wid - 1] : 0) + carry + sc - v;
    if (i < N) { un[i] = excl; rp[i] = excl; }
    __syncthreads();
    if (tid == 0) carry_s = carry + wsum[nw - 1];
    __syncthreads();
  }
  const int total = carry_s;
  if (tid == 0) {
    rp[N] = total;
    if (colidx && (long long)total > nnz_stride && overflow) atomicMax(overflow, total);
  }
  if (!colidx) return;

  // 5. fill.  Fast path (the graph's nnz fits the shared-memory stage): the row's matches go into the stage in cell order, then
  //    every match is written at its RANK among the row's matches (deg^2 compares, branch-free) — ascending columns without
  //    the 9-way merge, whose per-candidate 9 branches ran at 13.6 of 32 lanes and 491k warp-instructions per graph (ncu,
  //    round 2: 172 us for 256 graphs).
  int32_t* ci = colidx + (size_t)b * nnz_stride;
  float* vv = vals ? vals + (size_t)b * nnz_stride : nullptr;
  if (total <= stage_cap) {
    for (int i = tid; i < N; i += nt) {
      CellRanges R;
      ranges(cid[i], R);
      const float2 pi = sp[i];
      const int rs = un[i];
      int w = rs;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        for (int q = R.lo[k]; q < R.hi[k]; ++q) {
          const int j = lst[q];
          if (edge(pi, i, j)) stage[w++] = j;
        }
      const int d = w - rs;
      double isd_i = 0.0;
      if (NORM) isd_i = inv_sqrt_deg(d);
      for (int a = 0; a < d; ++a) {
        const int va = stage[rs + a];
        int r = 0;
        for (int c = 0; c < d; ++c) r += stage[rs + c] < va ? 1 : 0;
        const long long o = (long long)rs + r;
        if (o < nnz_stride) {
          ci[o] = va;
          if (vv) vv[o] = NORM ? (float)__dmul_rn(inv_sqrt_deg(aux[va]), isd_i) : 1.f;
        }
      }
    }
    return;
  }
  // general path: 9-way merge of the ascending cell lists
  for (int i = tid; i < N; i += nt) {
    const int c = cid[i];
    const int iy = c / gx, ix = c - iy * gx;
    int hd[9], en[9], hv[9];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int y2 = iy + d - 1, x2 = ix + e - 1;
        const bool ok = y2 >= 0 && y2 < gy && x2 >= 0 && x2 < gx;
        const int cc = ok ? y2 * gx + x2 : 0;
        hd[d * 3 + e] = ok ? cellstart[cc] : 0;
        en[d * 3 + e] = ok ? cellstart[cc + 1] : 0;
        hv[d * 3 + e] = hd[d * 3 + e] < en[d * 3 + e] ? lst[hd[d * 3 + e]] : 0x7fffffff;
      }
    const float2 pi = sp[i];
    long long w = un[i];
    double isd_i = 0.0;
    if (NORM) isd_i = inv_sqrt_deg(aux[i]);
    for (;;) {
      int best = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < 9; ++k) best = min(best, hv[k]);
      if (best == 0x7fffffff) break;
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (hv[k] == best) {   // node indices are unique across the lists: exactly one k matches
          ++hd[k];
          hv[k] = hd[k] < en[k] ? lst[hd[k]] : 0x7fffffff;
        }
      if (edge(pi, i, best)) {
        if (w < nnz_stride) {
          ci[w] = best;
          if (vv) vv[w] = NORM ? (float)__dmul_rn(inv_sqrt_deg(aux[best]), isd_i) : 1.f;
        }
        ++w;
      }
    }
  }
}

}  // namespace gfc

using namespace gfc;

// fp32 screening band around the squared-distance threshold (same band as set_thresholds in gfc_api.cu: an fp32
// squared distance differs from the fp64 one by < 2^-22 relative; 4e-6 on each side)
static void screen_band(double thr, float* lo, float* hi) {
  if (!(thr >= 0)) { *lo = -1.f; *hi = -1.f; return; }
  *lo = nextafterf((float)(thr * (1.0 - 4e-6)), -INFINITY);
  *hi = nextafterf((float)(thr * (1.0 + 4e-6)), INFINITY);
}

static int mode_threshold(int mode, double radius, double* thr, bool* norm) {
  switch (mode) {
    case GFC_GSO_BINARY_LE: *thr = squared_threshold(radius, true); *norm = false; return GFC_OK;
    case GFC_GSO_SYM_NORM_LT: *thr = squared_threshold(radius, false); *norm = true; return GFC_OK;
    case GFC_GSO_BINARY_LT: *thr = squared_threshold(radius, false); *norm = false; return GFC_OK;
    default: set_error("unknown GSO mode %d", mode); return GFC_ERR_BAD_ARG;
  }
}

// exported for the fused kernels in gfc_tile.cu
namespace gfc {
int gso_mode_threshold(int mode, double radius, double* thr, bool* norm) {
  return mode_threshold(mode, radius, thr, norm);
}
}

extern "C" int gfc_gso_build(const float* pos, int B, int N, double radius, int mode,
                             uint8_t* adj_out, float* S_out, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_gso_build: negative size B=%d N=%d", B, N);
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos != nullptr, GFC_ERR_BAD_ARG, "gfc_gso_build: pos is NULL");
  GFC_REQUIRE(adj_out || S_out, GFC_ERR_BAD_ARG, "gfc_gso_build: both outputs NULL");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  GFC_REQUIRE((long long)N * N <= 0x7fffffffLL / 2, GFC_ERR_UNSUPPORTED,
              "gfc_gso_build: N=%d too large for the dense builder (use gfc_csr_*)", N);
  long long per = (long long)N * N;
  int gpc = (int)(2048 / per);
  if (gpc < 1) gpc = 1;
  if (gpc > 64) gpc = 64;
  if (gpc > B) gpc = B;
  size_t smem = (size_t)gpc * N * 2 * sizeof(float) + (norm ? (size_t)gpc * N * sizeof(double) : 0);
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  GFC_REQUIRE(smem <= (size_t)di.smem_optin, GFC_ERR_UNSUPPORTED,
              "gfc_gso_build: N=%d needs %zu B shared memory (> %d)", N, smem, di.smem_optin);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(B, gpc);
  if (norm) {
    if (smem > 48 * 1024)
      GFC_CUDA_TRY(cudaFuncSetAttribute(gso_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gso_build_kernel<true><<<grid, 256, smem, st>>>(pos, B, N, thr, gpc, adj_out, S_out);
  } else {
    if (smem > 48 * 1024)
      GFC_CUDA_TRY(cudaFuncSetAttribute(gso_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gso_build_kernel<false><<<grid, 256, smem, st>>>(pos, B, N, thr, gpc, adj_out, S_out);
  }
  GFC_LAUNCH_CHECK("gso_build_kernel");
  return GFC_OK;
}

extern "C" int gfc_csr_count(const float* pos, int B, int N, double radius, int mode,
                             int32_t* deg, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_csr_count: negative size");
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos && deg, GFC_ERR_BAD_ARG, "gfc_csr_count: NULL pointer");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  float lo, hi;
  screen_band(thr, &lo, &hi);
  for (int b0 = 0; b0 < B; b0 += 65535) {   // gridDim.y limit: 65535 graphs per launch
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    csr_rows_kernel<false, false><<<dim3((unsigned)((N + 255) / 256), (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(
        pos + (size_t)b0 * N * 2, N, thr, lo, hi, deg + (size_t)b0 * N, nullptr, 0, nullptr, nullptr);
    GFC_LAUNCH_CHECK("csr_rows_kernel<count>");
  }
  return GFC_OK;
}

extern "C" int gfc_csr_scan(const int32_t* deg, int B, int N, int32_t* rowptr, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_csr_scan: negative size");
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(deg && rowptr, GFC_ERR_BAD_ARG, "gfc_csr_scan: NULL pointer");
  csr_scan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(deg, N, rowptr);
  GFC_LAUNCH_CHECK("csr_scan_kernel");
  return GFC_OK;
}

extern "C" int gfc_csr_fill(const float* pos, int B, int N, double radius, int mode,
                            const int32_t* rowptr, int64_t nnz_stride,
                            int32_t* colidx, float* vals, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0 && nnz_stride >= 0, GFC_ERR_BAD_ARG, "gfc_csr_fill: negative size");
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos && rowptr && colidx, GFC_ERR_BAD_ARG, "gfc_csr_fill: NULL pointer");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  float lo, hi;
  screen_band(thr, &lo, &hi);
  for (int b0 = 0; b0 < B; b0 += 65535) {   // gridDim.y limit: 65535 graphs per launch
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    const dim3 grid((unsigned)((N + 255) / 256), (unsigned)nb);
    const float* p0 = pos + (size_t)b0 * N * 2;
    const int32_t* rp0 = rowptr + (size_t)b0 * (N + 1);
    int32_t* ci0 = colidx + (size_t)b0 * nnz_stride;
    float* v0 = vals ? vals + (size_t)b0 * nnz_stride : nullptr;
    if (norm)
      csr_rows_kernel<true, true><<<grid, 256, 0, (cudaStream_t)stream>>>(p0, N, thr, lo, hi, nullptr, rp0, nnz_stride,
                                                                           ci0, v0);
    else
      csr_rows_kernel<true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p0, N, thr, lo, hi, nullptr, rp0, nnz_stride,
                                                                            ci0, v0);
    GFC_LAUNCH_CHECK("csr_rows_kernel<fill>");
  }
  return GFC_OK;
}

extern "C" int gfc_csr_build(const float* pos, int B, int N, double radius, int mode, int32_t* rowptr,
                             int64_t nnz_stride, int32_t* colidx, float* vals, int32_t* overflow, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0 && nnz_stride >= 0, GFC_ERR_BAD_ARG, "gfc_csr_build: negative size");
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(pos && rowptr, GFC_ERR_BAD_ARG, "gfc_csr_build: NULL pointer");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  float lo, hi;
  screen_band(thr, &lo, &hi);
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  const size_t smem_base = (size_t)N * (sizeof(float2) + 4 * sizeof(int)) + (size_t)(kCellMax + 1) * sizeof(int);
  GFC_REQUIRE(smem_base + 2048 <= (size_t)di.smem_optin, GFC_ERR_UNSUPPORTED,
              "gfc_csr_build: N=%d needs %zu B shared memory (> %d); use gfc_csr_count/scan/fill", N, smem_base, di.smem_optin);
  // the rest of the shared memory (the kernel runs one CTA per SM anyway: 50 registers x 1024 threads) stages the matches
  size_t stage_bytes = (size_t)di.smem_optin - 2048 - smem_base;
  if (colidx == nullptr) stage_bytes = 0;
  const size_t want = (size_t)N * 64 * sizeof(int);            // 64 neighbours per node on average is plenty
  if (stage_bytes > want) stage_bytes = want;
  const int stage_cap = (int)(stage_bytes / sizeof(int));
  const size_t smem = smem_base + (size_t)stage_cap * sizeof(int);
  // cell edge: 1.001 R (>= the largest distance the rule can accept, with room for the fp32 index rounding)
  const double r_acc = thr > 0 ? sqrt(thr) : 0.0;
  float cell0 = (float)(r_acc * 1.001);
  if (!(cell0 > 0.f)) cell0 = 0.f;   // no edges possible (or everything): one cell
  const int threads = N >= 1024 ? 1024 : (N > 512 ? 1024 : (N > 256 ? 512 : 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (norm) {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_build_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_build_fused_kernel<true><<<B, threads, smem, st>>>(pos, N, thr, lo, hi, cell0, rowptr, nnz_stride, colidx, vals, overflow, stage_cap);
  } else {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_build_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_build_fused_kernel<false><<<B, threads, smem, st>>>(pos, N, thr, lo, hi, cell0, rowptr, nnz_stride, colidx, vals, overflow, stage_cap);
  }
  GFC_LAUNCH_CHECK("csr_build_fused_kernel");
  return GFC_OK;
}
