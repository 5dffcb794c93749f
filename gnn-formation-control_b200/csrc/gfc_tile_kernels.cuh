// gfc_tile_kernels.cuh — device code of kernels (b) and (c): fused K-hop graph
// filter forward / backward for batches of small agent graphs (path A).
//
// What the reference does with ~8 ATen launches forward and ~20 backward
// (utils/graphUtils/graphML.py:2342-2366 + autograd, all in fp64) is done here
// in ONE persistent kernel per direction:
//   * a tile = `gpc` whole graphs = rows r = (graph j, node n) of a row-packed
//     matrix Z[r][k*G+g] kept in shared memory (z_k never touches HBM);
//   * the GSO tile comes either from dense S or is rebuilt from positions on
//     chip (fp32 screen + fp64 decision inside the rounding band: bit-identical to
//     scene.py:140-154 / multirobotsim_dcenlocal.py:306-315);
//   * the K-1 diffusion hops z_k = z_{k-1} S (graphML.py:2349-2352) run on the
//     FP32 pipes straight out of shared memory, skipping zero weights;
//   * the tap contraction y = Z H^T (graphML.py:2361-2362) runs on the tensor
//     cores as a 3xTF32 split product (fp32-equivalent accuracy) with the taps
//     pre-split and pre-swizzled into MMA B-fragment order;
//   * bias + activation are fused in the epilogue; y is written node-major
//     [B,N,F], which is the memory layout the reference returns.
// Backward recomputes Z, then dH += D^T Z, U = D H, Horner acc = acc S^T + U_k,
// dX = acc, db = colsum(D), with per-CTA partials reduced deterministically.
//
// TileCfg carries compile-time shapes (0 = runtime) so that the BASELINE shapes
// get shift/mask index math and fully unrolled inner loops.
#pragma once
#include "gfc_tile.cuh"

namespace gfc {

template <int N_, int G_, int F_, int K_, int THREADS_, int NB_>
struct TileCfg {
  static constexpr int sN = N_, sG = G_, sF = F_, sK = K_;
  static constexpr int kThreads = THREADS_, kWarps = THREADS_ / 32, kNB = NB_;
};

#define GFC_TILE_DIMS(CFG, p)                           \
  const int N = CFG::sN > 0 ? CFG::sN : (p).N;          \
  const int G = CFG::sG > 0 ? CFG::sG : (p).G;          \
  const int F = CFG::sF > 0 ? CFG::sF : (p).F;          \
  const int K = CFG::sK > 0 ? CFG::sK : (p).K;          \
  const int KG = K * G;                                 \
  (void)N; (void)G; (void)F; (void)K; (void)KG

// ---------------------------------------------------------------------------
// tap packing: B-fragment order, hi/lo split.
//   forward  (for_bwd=0): Bm[c][f] = h[f*KG+c], k-steps over c, n-tiles over f
//   backward (for_bwd=1): Bm[f][c] = h[f*KG+c], k-steps over f, n-tiles over c
// element p = (s*NT + nt)*32 + lane  ->  {hi(b0), hi(b1), lo(b0), lo(b1)}
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 pack_one(const float* __restrict__ h, int F, int KG,
                                           int for_bwd, int p) {
  const int lane = p & 31, q = p >> 5;
  const int g = lane >> 2, t = lane & 3;
  float b0, b1;
  if (!for_bwd) {
    const int NT = F >> 3;
    const int nt = q % NT, s = q / NT;
    const float* src = h + (size_t)(nt * 8 + g) * KG + s * 8 + t;
    b0 = __ldg(src);
    b1 = __ldg(src + 4);
  } else {
    const int NT = KG >> 3;
    const int nt = q % NT, s = q / NT;
    const float* src = h + (size_t)(s * 8 + t) * KG + nt * 8 + g;
    b0 = __ldg(src);
    b1 = __ldg(src + (size_t)4 * KG);
  }
  uint32_t h0, l0, h1, l1;
  split_tf32(b0, h0, l0);
  split_tf32(b1, h1, l1);
  return make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0),
                     __uint_as_float(l1));
}

// ---------------------------------------------------------------------------
// tile loaders
// ---------------------------------------------------------------------------
// x[b0 .. b0+gcount) [G][N]  ->  Z[(j*N+n)*ldz + g]   (4x4 register transpose when N % 4 == 0)
template <typename CFG>
__device__ __forceinline__ void load_x_tile(float* __restrict__ Zs, const float* __restrict__ x,
                                            const TilePlan& p, int b0, int gcount, int ldz, int vec_ok) {
  GFC_TILE_DIMS(CFG, p);
  const int tid = threadIdx.x;
  const int GN = G * N;
  const float* src = x + (size_t)b0 * GN;
  if (vec_ok && N == 1 && (G & 3) == 0) {   // rows are already contiguous: 128-bit copies
    const int G4 = G >> 2, total = gcount * G4;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    constexpr int U = 8;   // loads in flight per thread: the tile load is the latency chain of this mode
    for (int q0 = tid; q0 < total; q0 += U * CFG::kThreads) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u * CFG::kThreads;
        if (q < total) v[u] = __ldg(s4 + q);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u * CFG::kThreads;
        if (q < total) {
          const int j = q / G4, g4 = q - j * G4;
          *reinterpret_cast<float4*>(Zs + (size_t)j * ldz + g4 * 4) = v[u];
        }
      }
    }
  } else if (vec_ok && (N & 3) == 0) {
    const int N4 = N >> 2, G4 = G >> 2;
    const int per_graph = N4 * G4;
    const int total = gcount * per_graph;
    for (int q = tid; q < total; q += CFG::kThreads) {
      const int j = q / per_graph, rem = q - j * per_graph;
      const int gb = rem / N4, nb4 = rem - gb * N4;
      const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)j * GN + (size_t)(gb * 4) * N + nb4 * 4);
      const float4 v0 = __ldg(s4);
      const float4 v1 = __ldg(s4 + N4);
      const float4 v2 = __ldg(s4 + 2 * N4);
      const float4 v3 = __ldg(s4 + 3 * N4);
      float* dst = Zs + (size_t)(j * N + nb4 * 4) * ldz + gb * 4;
      *reinterpret_cast<float4*>(dst) = make_float4(v0.x, v1.x, v2.x, v3.x);
      *reinterpret_cast<float4*>(dst + ldz) = make_float4(v0.y, v1.y, v2.y, v3.y);
      *reinterpret_cast<float4*>(dst + 2 * ldz) = make_float4(v0.z, v1.z, v2.z, v3.z);
      *reinterpret_cast<float4*>(dst + 3 * ldz) = make_float4(v0.w, v1.w, v2.w, v3.w);
    }
  } else {
    const int total = gcount * GN;
    for (int i = tid; i < total; i += CFG::kThreads) {
      const int j = i / GN, rem = i - j * GN;
      const int g = rem / N, n = rem - g * N;
      Zs[(size_t)(j * N + n) * ldz + g] = __ldg(src + i);
    }
  }
}

// dX[b0 .. b0+gcount) [G][N]  <-  Z[(j*N+n)*ldz + g]
template <typename CFG>
__device__ __forceinline__ void store_dx_tile(const float* __restrict__ Zs, float* __restrict__ dX,
                                              const TilePlan& p, int b0, int gcount, int ldz, int vec_ok) {
  GFC_TILE_DIMS(CFG, p);
  const int tid = threadIdx.x;
  const int GN = G * N;
  float* dst = dX + (size_t)b0 * GN;
  if (vec_ok && N == 1 && (G & 3) == 0) {
    const int G4 = G >> 2, total = gcount * G4;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int q = tid; q < total; q += CFG::kThreads) {
      const int j = q / G4, g4 = q - j * G4;
      d4[q] = *reinterpret_cast<const float4*>(Zs + (size_t)j * ldz + g4 * 4);
    }
  } else if (vec_ok && (N & 3) == 0) {
    const int N4 = N >> 2, G4 = G >> 2;
    const int per_graph = N4 * G4;
    const int total = gcount * per_graph;
    for (int q = tid; q < total; q += CFG::kThreads) {
      const int j = q / per_graph, rem = q - j * per_graph;
      const int gb = rem / N4, nb4 = rem - gb * N4;
      const float* s = Zs + (size_t)(j * N + nb4 * 4) * ldz + gb * 4;
      const float4 v0 = *reinterpret_cast<const float4*>(s);
      const float4 v1 = *reinterpret_cast<const float4*>(s + ldz);
      const float4 v2 = *reinterpret_cast<const float4*>(s + 2 * ldz);
      const float4 v3 = *reinterpret_cast<const float4*>(s + 3 * ldz);
      float4* d4 = reinterpret_cast<float4*>(dst + (size_t)j * GN + (size_t)(gb * 4) * N + nb4 * 4);
      d4[0] = make_float4(v0.x, v1.x, v2.x, v3.x);
      d4[N4] = make_float4(v0.y, v1.y, v2.y, v3.y);
      d4[2 * N4] = make_float4(v0.z, v1.z, v2.z, v3.z);
      d4[3 * N4] = make_float4(v0.w, v1.w, v2.w, v3.w);
    }
  } else {
    const int total = gcount * GN;
    for (int i = tid; i < total; i += CFG::kThreads) {
      const int j = i / GN, rem = i - j * GN;
      const int gg = rem / N, n = rem - gg * N;
      dst[i] = Zs[(size_t)(j * N + n) * ldz + gg];
    }
  }
}

// adjacency decision for one pair: fp32 screen, fp64 rule inside the rounding band.  The exact rule is
// kept out of line so that the (rare) band case is a real branch: if-converted, its fp64 and conversion
// instructions ran for every pair and dominated the GSO phase (profiles/r1).
static __device__ __noinline__ bool pair_adjacent_exact(float xi, float yi, float xj, float yj, double thr) {
  return sqdist64(xi, yi, xj, yj) <= thr;
}
__device__ __forceinline__ bool pair_adjacent(float xi, float yi, float xj, float yj, const TileArgs& a) {
  const float dx = xi - xj, dy = yi - yj;
  const float s = fmaf(dx, dx, dy * dy);
  if (s < a.thr_lo) return true;
  if (s > a.thr_hi) return false;
  return pair_adjacent_exact(xi, yi, xj, yj, a.thr);
}

// GSO tile Ss[j][m][n] from dense S (GSRC_DENSE) or rebuilt from positions.
template <typename CFG, int GSRC>
__device__ __forceinline__ void load_gso_tile(float* __restrict__ Ss, float* __restrict__ sp,
                                              double* __restrict__ isd, const TileArgs& a,
                                              int b0, int gcount) {
  GFC_TILE_DIMS(CFG, a.p);
  const int tid = threadIdx.x;
  const int NN = N * N;
  if (GSRC == GSRC_DENSE) {
    const int total = gcount * NN;
    if (a.s_bstride == 0) {   // one GSO shared by the whole batch (GraphFilter, graphML.py:1111): every slot gets the same matrix
      for (int i = tid; i < total; i += CFG::kThreads) Ss[i] = __ldg(a.S + i % NN);
    } else {
      const float* src = a.S + (size_t)b0 * NN;
      if (a.vec_ok && (NN & 3) == 0) {
        const float4* src4 = reinterpret_cast<const float4*>(src);
        float4* dst4 = reinterpret_cast<float4*>(Ss);
        for (int i = tid; i < (total >> 2); i += CFG::kThreads) dst4[i] = __ldg(src4 + i);
      } else {
        for (int i = tid; i < total; i += CFG::kThreads) Ss[i] = __ldg(src + i);
      }
    }
  } else {
    const int nn = gcount * N;
    const float* gp = a.pos + (size_t)b0 * N * 2;
    for (int i = tid; i < nn * 2; i += CFG::kThreads) sp[i] = __ldg(gp + i);
    __syncthreads();
    // the rule is symmetric: decide the pairs m < n once and mirror them
    const int total = nn * N;
    for (int o = tid; o < total; o += CFG::kThreads) {
      const int r = o / N, n2 = o - r * N;
      const int j = r / N, m = r - j * N;
      if (m < n2) {
        const int q = j * N + n2;
        const float v = pair_adjacent(sp[2 * r], sp[2 * r + 1], sp[2 * q], sp[2 * q + 1], a) ? 1.f : 0.f;
        Ss[o] = v;
        Ss[(size_t)q * N + m] = v;
      } else if (m == n2) {
        Ss[o] = 0.f;
      }
    }
    if (a.norm) {
      __syncthreads();
      for (int r = tid; r < nn; r += CFG::kThreads) {
        float deg = 0.f;
        for (int m = 0; m < N; ++m) deg += Ss[(size_t)r * N + m];
        isd[r] = inv_sqrt_deg((int)deg);
      }
      __syncthreads();
      for (int o = tid; o < total; o += CFG::kThreads) {
        const int r = o / N, n2 = o - r * N;
        const int j = r / N;
        if (Ss[o] != 0.f) Ss[o] = (float)__dmul_rn(isd[r], isd[j * N + n2]);
      }
    }
  }
}

// One diffusion hop in shared memory.
//   TRANSPOSED = false: Z[r][dst] = sum_m S_j[m][n] Z[(j,m)][src]          (z_k = z_{k-1} S)
//   TRANSPOSED = true : Z[r][dst] += sum_m S_j[n][m] Z[(j,m)][src]         (acc S^T + U_k)
template <typename CFG, bool TRANSPOSED>
__device__ __forceinline__ void hop_tile(float* __restrict__ Zs, const float* __restrict__ Ss,
                                         const TilePlan& p, int rows_used, int ldz, int src_col, int dst_col) {
  GFC_TILE_DIMS(CFG, p);
  const int G4 = G >> 2;
  const int total = rows_used * G4;
  for (int idx = threadIdx.x; idx < total; idx += CFG::kThreads) {
    const int r = idx / G4, g4 = idx - r * G4;
    const int j = r / N, n = r - j * N;
    const float* sw = Ss + (size_t)j * N * N + (TRANSPOSED ? n * N : n);
    const int sstride = TRANSPOSED ? 1 : N;
    const float* zin = Zs + (size_t)(j * N) * ldz + src_col + (g4 << 2);
    float* zout = Zs + (size_t)r * ldz + dst_col + (g4 << 2);
    float4 acc = TRANSPOSED ? *reinterpret_cast<const float4*>(zout) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (CFG::sN > 0 && CFG::sN <= 16) {
#pragma unroll
      for (int m = 0; m < (CFG::sN > 0 ? CFG::sN : 1); ++m) {
        const float w = sw[m * sstride];
        if (w != 0.f) {
          const float4 z = *reinterpret_cast<const float4*>(zin + (size_t)m * ldz);
          acc.x = fmaf(w, z.x, acc.x); acc.y = fmaf(w, z.y, acc.y);
          acc.z = fmaf(w, z.z, acc.z); acc.w = fmaf(w, z.w, acc.w);
        }
      }
    } else {
      for (int m = 0; m < N; ++m) {
        const float w = sw[m * sstride];
        if (w != 0.f) {
          const float4 z = *reinterpret_cast<const float4*>(zin + (size_t)m * ldz);
          acc.x = fmaf(w, z.x, acc.x); acc.y = fmaf(w, z.y, acc.y);
          acc.z = fmaf(w, z.z, acc.z); acc.w = fmaf(w, z.w, acc.w);
        }
      }
    }
    *reinterpret_cast<float4*>(zout) = acc;
  }
}

// Compact neighbour lists of the GSO tile: for row r = (j, n) the indices m with a non-zero
// weight, column-wise (S[m][n], forward hops) or row-wise (S[n][m], Horner).  u8 indices
// [rows][N] followed by u8 counts [rows].
template <typename CFG, bool TRANSPOSED>
__device__ __forceinline__ void build_lists(unsigned char* __restrict__ lst, const float* __restrict__ Ss,
                                            const TilePlan& p, int rows_used) {
  GFC_TILE_DIMS(CFG, p);
  unsigned char* cnt = lst + (size_t)p.rows * N;
  for (int r = threadIdx.x; r < rows_used; r += CFG::kThreads) {
    const int j = r / N, n = r - j * N;
    const float* sw = Ss + (size_t)j * N * N + (TRANSPOSED ? n * N : n);
    const int sstride = TRANSPOSED ? 1 : N;
    unsigned char* out = lst + (size_t)r * N;
    int c = 0;
    for (int m = 0; m < N; ++m)
      if (sw[m * sstride] != 0.f) out[c++] = (unsigned char)m;
    cnt[r] = (unsigned char)c;
  }
}

template <typename CFG, bool TRANSPOSED>
__device__ __forceinline__ void hop_tile_lists(float* __restrict__ Zs, const float* __restrict__ Ss,
                                               const unsigned char* __restrict__ lst, const TilePlan& p,
                                               int rows_used, int ldz, int src_col, int dst_col) {
  GFC_TILE_DIMS(CFG, p);
  const int G4 = G >> 2;
  const int total = rows_used * G4;
  const unsigned char* cnt = lst + (size_t)p.rows * N;
  for (int idx = threadIdx.x; idx < total; idx += CFG::kThreads) {
    const int r = idx / G4, g4 = idx - r * G4;
    const int j = r / N, n = r - j * N;
    const float* sw = Ss + (size_t)j * N * N + (TRANSPOSED ? n * N : n);
    const int sstride = TRANSPOSED ? 1 : N;
    const float* zin = Zs + (size_t)(j * N) * ldz + src_col + (g4 << 2);
    float* zout = Zs + (size_t)r * ldz + dst_col + (g4 << 2);
    float4 acc = TRANSPOSED ? *reinterpret_cast<const float4*>(zout) : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned char* li = lst + (size_t)r * N;
    const int c = cnt[r];
    for (int i = 0; i < c; ++i) {
      const int m = li[i];
      const float w = sw[m * sstride];
      const float4 z = *reinterpret_cast<const float4*>(zin + (size_t)m * ldz);
      acc.x = fmaf(w, z.x, acc.x); acc.y = fmaf(w, z.y, acc.y);
      acc.z = fmaf(w, z.z, acc.z); acc.w = fmaf(w, z.w, acc.w);
    }
    *reinterpret_cast<float4*>(zout) = acc;
  }
}

// ---------------------------------------------------------------------------
// Register-column fast path (static small N, one (graph j, feature g) column per thread):
// the x column is loaded from global into registers, all K-1 hops run back to back in
// registers (dense N x N FMAs, S read as shared-memory broadcasts), and every state is
// written once into its Z slot.  No barrier between hops.
// ---------------------------------------------------------------------------
template <typename CFG>
struct FastPath {
  static constexpr bool kEnabled = (CFG::sN > 0 && CFG::sN <= 16 && CFG::sG > 0 && CFG::sK > 0 &&
                                    (128 / (CFG::sN > 0 ? CFG::sN : 1)) * CFG::sG <= CFG::kThreads);
};

template <typename CFG>
__device__ __forceinline__ void load_x_column(float (&xr)[CFG::sN > 0 ? CFG::sN : 1], const float* __restrict__ x,
                                              int b0, int j, int g, int vec_ok) {
  constexpr int N = CFG::sN > 0 ? CFG::sN : 1, G = CFG::sG;
  const float* src = x + ((size_t)(b0 + j) * G + g) * N;
  if ((N & 3) == 0 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
      xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int n = 0; n < N; ++n) xr[n] = __ldg(src + n);
  }
}

// z_0 = xr; z_k = z_{k-1} S_j for k = 1..K-1; all states written to Z[(j,n)][k*G+g]
template <typename CFG>
__device__ __forceinline__ void hops_from_column(float (&z)[CFG::sN > 0 ? CFG::sN : 1], float* __restrict__ Zs,
                                                 const float* __restrict__ Sj, int j, int g, int ldz) {
  constexpr int N = CFG::sN > 0 ? CFG::sN : 1, G = CFG::sG, K = CFG::sK;
  float* zc = Zs + (size_t)(j * N) * ldz + g;
#pragma unroll
  for (int n = 0; n < N; ++n) zc[(size_t)n * ldz] = z[n];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    float zn[N];
#pragma unroll
    for (int n = 0; n < N; ++n) zn[n] = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float zm = z[m];
      if ((N & 3) == 0) {  // 128-bit broadcast loads of row m of S_j
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
          const float4 sv = *reinterpret_cast<const float4*>(Sj + m * N + 4 * q);
          zn[4 * q] = fmaf(sv.x, zm, zn[4 * q]);
          zn[4 * q + 1] = fmaf(sv.y, zm, zn[4 * q + 1]);
          zn[4 * q + 2] = fmaf(sv.z, zm, zn[4 * q + 2]);
          zn[4 * q + 3] = fmaf(sv.w, zm, zn[4 * q + 3]);
        }
      } else {
#pragma unroll
        for (int n = 0; n < N; ++n) zn[n] = fmaf(Sj[m * N + n], zm, zn[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < N; ++n) { z[n] = zn[n]; zc[(size_t)n * ldz + k * G] = zn[n]; }
  }
}

// Horner over the U slots: acc = U_{K-1}; acc = acc S_j^T + U_k; result (= dX column) in acc
template <typename CFG>
__device__ __forceinline__ void horner_from_column(float (&acc)[CFG::sN > 0 ? CFG::sN : 1], const float* __restrict__ Zs,
                                                   const float* __restrict__ Sj, int j, int g, int ldz) {
  constexpr int N = CFG::sN > 0 ? CFG::sN : 1, G = CFG::sG, K = CFG::sK;
  const float* zc = Zs + (size_t)(j * N) * ldz + g;
#pragma unroll
  for (int n = 0; n < N; ++n) acc[n] = zc[(size_t)n * ldz + (K - 1) * G];
#pragma unroll
  for (int k = K - 2; k >= 0; --k) {
    float an[N];
#pragma unroll
    for (int n = 0; n < N; ++n) an[n] = zc[(size_t)n * ldz + k * G];
#pragma unroll
    for (int n = 0; n < N; ++n) {   // an[n] += sum_m S[n][m] acc[m], row n of S_j as 128-bit broadcasts
      float sacc = an[n];
      if ((N & 3) == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
          const float4 sv = *reinterpret_cast<const float4*>(Sj + n * N + 4 * q);
          sacc = fmaf(sv.x, acc[4 * q], sacc);
          sacc = fmaf(sv.y, acc[4 * q + 1], sacc);
          sacc = fmaf(sv.z, acc[4 * q + 2], sacc);
          sacc = fmaf(sv.w, acc[4 * q + 3], sacc);
        }
      } else {
#pragma unroll
        for (int m = 0; m < N; ++m) sacc = fmaf(Sj[n * N + m], acc[m], sacc);
      }
      an[n] = sacc;
    }
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = an[n];
  }
}

// GSO tile straight from global positions (no staging, one barrier less); binary modes and
// sym-norm (two extra barriers, taken uniformly by the whole CTA).
template <typename CFG>
__device__ __forceinline__ void gso_tile_from_global(float* __restrict__ Ss, double* __restrict__ isd,
                                                     const TileArgs& a, int b0, int gcount) {
  constexpr int N = CFG::sN > 0 ? CFG::sN : 1;
  const int nn = gcount * N, total = nn * N;
  const float2* gp = reinterpret_cast<const float2*>(a.pos) + (size_t)b0 * N;
  for (int o = threadIdx.x; o < total; o += CFG::kThreads) {
    const int r = o / N, n2 = o - r * N;
    const int j = r / N, m = r - j * N;
    if (m < n2) {
      const int q = j * N + n2;
      const float2 pi = __ldg(gp + r), pj = __ldg(gp + q);
      const float v = pair_adjacent(pi.x, pi.y, pj.x, pj.y, a) ? 1.f : 0.f;
      Ss[o] = v;
      Ss[(size_t)q * N + m] = v;
    } else if (m == n2) {
      Ss[o] = 0.f;
    }
  }
  if (a.norm) {
    __syncthreads();
    for (int r = threadIdx.x; r < nn; r += CFG::kThreads) {
      float deg = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) deg += Ss[(size_t)r * N + m];
      isd[r] = inv_sqrt_deg((int)deg);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < total; o += CFG::kThreads) {
      const int r = o / N, n2 = o - r * N;
      const int j = r / N;
      if (Ss[o] != 0.f) Ss[o] = (float)__dmul_rn(isd[r], isd[j * N + n2]);
    }
  }
}

// debug: thread 0 stamps the SM clock at phase boundaries of the CTA's first tile
#define GFC_STAMP(a, slot)                                                                 \
  do {                                                                                     \
    if ((a).dbg_clk && threadIdx.x == 0 && (slot) < 16) (a).dbg_clk[(size_t)blockIdx.x * 16 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GFC_STAMP_NS(a, slot)                                                              \
  do {                                                                                     \
    if ((a).dbg_clk && threadIdx.x == 0) (a).dbg_clk[(size_t)blockIdx.x * 16 + (slot)] = global_ns(); \
  } while (0)

template <int THREADS>
__device__ __forceinline__ void zero_floats(float* p, int n) {
  for (int i = threadIdx.x; i < n; i += THREADS) p[i] = 0.f;
}

// three-term (or single-pass) tensor-core product accumulate
__device__ __forceinline__ void mma3(float (&acc)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                     uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1, bool single) {
  if (!single) {
    mma_tf32(acc, alo, bh0, bh1);
    mma_tf32(acc, ahi, bl0, bl1);
  }
  mma_tf32(acc, ahi, bh0, bh1);
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <typename CFG, int GSRC, bool HSMEM>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kThreads >= 512 ? 2 : 1)
tile_fwd_kernel(const TileArgs a) {
  extern __shared__ __align__(16) float smem[];
  const TilePlan& p = a.p;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);
  GFC_TILE_DIMS(CFG, p);
  constexpr int NB = CFG::kNB;
  float* Zs = smem + p.off_z;
  float* Ss = smem + p.off_s;
  float* sp = smem + p.off_pos;
  double* isd = reinterpret_cast<double*>(smem + p.off_isd);
  float4* Hs = reinterpret_cast<float4*>(smem + p.off_h);
  unsigned char* lst_c = reinterpret_cast<unsigned char*>(smem + p.off_nbr);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ldz = KG + 4;
  const bool single = a.single_pass != 0;

  // MMA pad rows (rows..rpad) are never written by a tile: clear them once
  if (p.rpad > p.rows) zero_floats<CFG::kThreads>(Zs + (size_t)p.rows * ldz, (p.rpad - p.rows) * ldz);
  if (HSMEM) {
    const int total = (KG * F) >> 1;
    for (int q = tid; q < total; q += CFG::kThreads) Hs[q] = pack_one(a.h, F, KG, 0, q);
  }
  const float4* HP = HSMEM ? Hs : a.hpack;
  GFC_STAMP(a, 0);

  const int MT = p.rpad >> 4, NT = F >> 3, KS = KG >> 3;
  const int ngroups = (NT + NB - 1) / NB;
  const int ntasks = MT * ngroups;

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int b0 = tile * p.gpc;
    const int gcount = min(p.gpc, p.B - b0);
    const int rows_used = gcount * N;
    if (gcount < p.gpc) {  // tail tile: clear the rows no graph maps to
      for (int i = tid; i < (p.rows - rows_used) * ldz; i += CFG::kThreads) Zs[(size_t)rows_used * ldz + i] = 0.f;
    }
    if constexpr (FastPath<CFG>::kEnabled) {
      // one (graph, feature) column per thread: x -> registers while the GSO tile is rebuilt
      const int j = tid / G, gcol = tid - j * G;
      const bool has_col = tid < gcount * G;
      float zcol[CFG::sN > 0 ? CFG::sN : 1];
      if (has_col) load_x_column<CFG>(zcol, a.x, b0, j, gcol, a.vec_ok);
      if (GSRC == GSRC_POS) gso_tile_from_global<CFG>(Ss, isd, a, b0, gcount);
      else load_gso_tile<CFG, GSRC>(Ss, sp, isd, a, b0, gcount);
      __syncthreads();
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 1);
      if (has_col) hops_from_column<CFG>(zcol, Zs, Ss + (size_t)j * N * N, j, gcol, ldz);
      __syncthreads();
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 2);
    } else {
      load_x_tile<CFG>(Zs, a.x, p, b0, gcount, ldz, a.vec_ok);
      load_gso_tile<CFG, GSRC>(Ss, sp, isd, a, b0, gcount);
      __syncthreads();
      if (p.use_lists && K > 1) {
        build_lists<CFG, false>(lst_c, Ss, p, rows_used);
        __syncthreads();
      }
      for (int k = 1; k < K; ++k) {
        if (p.use_lists) hop_tile_lists<CFG, false>(Zs, Ss, lst_c, p, rows_used, ldz, (k - 1) * G, k * G);
        else hop_tile<CFG, false>(Zs, Ss, p, rows_used, ldz, (k - 1) * G, k * G);
        __syncthreads();
      }
    }
    // ---- tap contraction on tensor cores: Y[rpad x F] = Z[rpad x KG] * Hm[KG x F]
    for (int task = warp; task < ntasks; task += CFG::kWarps) {
      const int mt = task / ngroups, ng = task - mt * ngroups;
      const int nt0 = ng * NB;
      const int nbc = min(NB, NT - nt0);
      float acc[NB][4];
#pragma unroll
      for (int i = 0; i < NB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const float* za = Zs + (size_t)(mt * 16 + g) * ldz + t;
      const float4* hp = HP + (size_t)nt0 * 32 + lane;
#pragma unroll 4
      for (int s = 0; s < KS; ++s) {
        uint32_t ahi[4], alo[4];
        split_tf32(za[s * 8], ahi[0], alo[0]);
        split_tf32(za[8 * ldz + s * 8], ahi[1], alo[1]);
        split_tf32(za[s * 8 + 4], ahi[2], alo[2]);
        split_tf32(za[8 * ldz + s * 8 + 4], ahi[3], alo[3]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (nb < nbc) {
            const float4 b = hp[((size_t)s * NT + nb) * 32];
            mma3(acc[nb], ahi, alo, __float_as_uint(b.x), __float_as_uint(b.y),
                 __float_as_uint(b.z), __float_as_uint(b.w), single);
          }
        }
      }
      // ---- epilogue: bias + activation, node-major store y[(b0*N + r)*F + f]
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      float* yrow0 = a.y + ((size_t)b0 * N + r0) * F;
      float* yrow1 = a.y + ((size_t)b0 * N + r1) * F;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (nb < nbc) {
          const int f0 = (nt0 + nb) * 8 + 2 * t;
          const float bb0 = a.bias ? __ldg(a.bias + f0) : 0.f;
          const float bb1 = a.bias ? __ldg(a.bias + f0 + 1) : 0.f;
          if (r0 < rows_used) {
            float2 v = make_float2(apply_act(acc[nb][0] + bb0, a.act, a.slope),
                                   apply_act(acc[nb][1] + bb1, a.act, a.slope));
            *reinterpret_cast<float2*>(yrow0 + f0) = v;
          }
          if (r1 < rows_used) {
            float2 v = make_float2(apply_act(acc[nb][2] + bb0, a.act, a.slope),
                                   apply_act(acc[nb][3] + bb1, a.act, a.slope));
            *reinterpret_cast<float2*>(yrow1 + f0) = v;
          }
        }
      }
    }
    __syncthreads();  // Z / S are overwritten by the next tile
    if (tile == (int)blockIdx.x) GFC_STAMP(a, 3);
  }
  GFC_STAMP_NS(a, 9);
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
template <typename CFG, int GSRC, bool HSMEM, bool ACC>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kThreads >= 512 ? 2 : 1)
tile_bwd_kernel(const TileArgs a) {
  extern __shared__ __align__(16) float smem[];
  const TilePlan& p = a.p;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);
  GFC_TILE_DIMS(CFG, p);
  constexpr int NB = CFG::kNB;
  float* Zs = smem + p.off_z;
  float* Ss = smem + p.off_s;
  float* Ds = smem + p.off_d;
  float* sp = smem + p.off_pos;
  double* isd = reinterpret_cast<double*>(smem + p.off_isd);
  float4* Hs = reinterpret_cast<float4*>(smem + p.off_h);
  float* dbs = Ds + (size_t)p.rpad * (F + 4);  // [F] running bias gradient of this CTA
  unsigned char* lst_c = reinterpret_cast<unsigned char*>(smem + p.off_nbr);
  unsigned char* lst_r = lst_c + ((((size_t)p.rows * N + p.rows + 3) >> 2) + 3 & ~(size_t)3) * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ldz = KG + 8, ldd = F + 4;
  const bool single = a.single_pass != 0;
  const bool want_dx = a.dX != nullptr, want_dh = a.dHp != nullptr, want_db = a.dbp != nullptr;

  if (p.rpad > p.rows) {
    zero_floats<CFG::kThreads>(Zs + (size_t)p.rows * ldz, (p.rpad - p.rows) * ldz);
    zero_floats<CFG::kThreads>(Ds + (size_t)p.rows * ldd, (p.rpad - p.rows) * ldd);
  }
  zero_floats<CFG::kThreads>(dbs, F);
  if (HSMEM && want_dx) {
    const int total = (KG * F) >> 1;
    for (int q = tid; q < total; q += CFG::kThreads) Hs[q] = pack_one(a.h, F, KG, 1, q);
  }
  const float4* HP = HSMEM ? Hs : a.hpack;
  __syncthreads();
  GFC_STAMP(a, 0);

  // dH task geometry: M = F, N = KG, Kdim = rows
  const int MTd = F >> 4, NTd = KG >> 3, KSd = p.rpad >> 3;
  const int nbd = p.nb_dh;
  const int ngroups_d = (NTd + nbd - 1) / nbd;
  const int ntasks_d = MTd * ngroups_d;
  // U task geometry: M = rows, N = KG, Kdim = F
  const int MTu = p.rpad >> 4, NTu = KG >> 3, KSu = F >> 3;
  const int ngroups_u = (NTu + NB - 1) / NB;
  const int ntasks_u = MTu * ngroups_u;

  float accH[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) accH[i][0] = accH[i][1] = accH[i][2] = accH[i][3] = 0.f;
  float* dHpart = want_dh ? a.dHp + (size_t)(ACC ? blockIdx.x : (blockIdx.x % p.nparts)) * F * KG : nullptr;

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int b0 = tile * p.gpc;
    const int gcount = min(p.gpc, p.B - b0);
    const int rows_used = gcount * N;
    if (gcount < p.gpc) {
      for (int i = tid; i < (p.rows - rows_used) * ldz; i += CFG::kThreads) Zs[(size_t)rows_used * ldz + i] = 0.f;
      for (int i = tid; i < (p.rows - rows_used) * ldd; i += CFG::kThreads) Ds[(size_t)rows_used * ldd + i] = 0.f;
    }
    // ---- loads: D = dY * act'(y), x -> Z_0, GSO tile
    {
      const int total = rows_used * F;
      const float* dsrc = a.dY + (size_t)b0 * N * F;
      const float* ysrc = (a.act != GFC_ACT_NONE) ? a.yout + (size_t)b0 * N * F : nullptr;
      if (a.vec_ok) {
        const float4* d4 = reinterpret_cast<const float4*>(dsrc);
        const float4* y4 = reinterpret_cast<const float4*>(ysrc);
        constexpr int U = CFG::kThreads >= 512 ? 1 : 4;   // (dY, y) pairs in flight per thread (64-register CTAs: 1)
        const int total4 = total >> 2;
        for (int q0 = tid; q0 < total4; q0 += U * CFG::kThreads) {
          float4 v[U], yo[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i4 = q0 + u * CFG::kThreads;
            if (i4 < total4) {
              v[u] = __ldg(d4 + i4);
              if (ysrc) yo[u] = __ldg(y4 + i4);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i4 = q0 + u * CFG::kThreads;
            if (i4 < total4) {
              if (ysrc) {
                v[u].x = act_grad(v[u].x, yo[u].x, a.act, a.slope);
                v[u].y = act_grad(v[u].y, yo[u].y, a.act, a.slope);
                v[u].z = act_grad(v[u].z, yo[u].z, a.act, a.slope);
                v[u].w = act_grad(v[u].w, yo[u].w, a.act, a.slope);
              }
              const int i = i4 << 2;
              const int r = i / F, f = i - r * F;
              *reinterpret_cast<float4*>(Ds + (size_t)r * ldd + f) = v[u];
            }
          }
        }
      } else {
        for (int i = tid; i < total; i += CFG::kThreads) {
          float v = __ldg(dsrc + i);
          if (ysrc) v = act_grad(v, __ldg(ysrc + i), a.act, a.slope);
          const int r = i / F, f = i - r * F;
          Ds[(size_t)r * ldd + f] = v;
        }
      }
    }
    constexpr bool kFast = FastPath<CFG>::kEnabled;
    const int jcol = kFast ? tid / G : 0, gcol = kFast ? tid - jcol * G : 0;
    const bool has_col = kFast && tid < gcount * G;
    float zcol[CFG::sN > 0 ? CFG::sN : 1];
    if constexpr (kFast) {
      if (want_dh && has_col) load_x_column<CFG>(zcol, a.x, b0, jcol, gcol, a.vec_ok);
      if (GSRC == GSRC_POS) gso_tile_from_global<CFG>(Ss, isd, a, b0, gcount);
      else load_gso_tile<CFG, GSRC>(Ss, sp, isd, a, b0, gcount);
      __syncthreads();
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 1);
    } else {
      if (want_dh) load_x_tile<CFG>(Zs, a.x, p, b0, gcount, ldz, a.vec_ok);
      load_gso_tile<CFG, GSRC>(Ss, sp, isd, a, b0, gcount);
      __syncthreads();
      if (p.use_lists && K > 1) {
        if (want_dh) build_lists<CFG, false>(lst_c, Ss, p, rows_used);
        if (want_dx) build_lists<CFG, true>(lst_r, Ss, p, rows_used);
        __syncthreads();
      }
    }

    // ---- db += column sums of D (all threads: thread -> column f, row class)
    if (want_db) {
      const int groups = CFG::kThreads / F;  // F <= kThreads is guaranteed by the plan
      if (groups > 0) {
        const int f = tid % F, grp = tid / F;
        if (grp < groups) {
          float s = 0.f;
          for (int r = grp; r < rows_used; r += groups) s += Ds[(size_t)r * ldd + f];
          atomicAdd(dbs + f, s);
        }
      }
    }

    if (want_dh) {
      // ---- recompute the diffusion states (graphML.py:2349-2352)
      if constexpr (kFast) {
        if (has_col) hops_from_column<CFG>(zcol, Zs, Ss + (size_t)jcol * N * N, jcol, gcol, ldz);
        __syncthreads();
        if (tile == (int)blockIdx.x) GFC_STAMP(a, 2);
      } else {
        for (int k = 1; k < K; ++k) {
          if (p.use_lists) hop_tile_lists<CFG, false>(Zs, Ss, lst_c, p, rows_used, ldz, (k - 1) * G, k * G);
          else hop_tile<CFG, false>(Zs, Ss, p, rows_used, ldz, (k - 1) * G, k * G);
          __syncthreads();
        }
      }
      // ---- dH[f][c] += sum_r D[r][f] Z[r][c]
      for (int task = warp; task < ntasks_d; task += CFG::kWarps) {
        const int mt = task / ngroups_d, ng = task - mt * ngroups_d;
        const int nt0 = ng * nbd;
        const int nbc = min(nbd, NTd - nt0);
        if (!ACC) {
#pragma unroll
          for (int i = 0; i < NB; ++i) accH[i][0] = accH[i][1] = accH[i][2] = accH[i][3] = 0.f;
        }
        const float* dcol = Ds + (size_t)t * ldd + mt * 16 + g;
        const float* zb = Zs + (size_t)t * ldz + nt0 * 8 + g;
#pragma unroll 4
        for (int s = 0; s < KSd; ++s) {
          const float* dc = dcol + (size_t)s * 8 * ldd;
          uint32_t ahi[4], alo[4];
          split_tf32(dc[0], ahi[0], alo[0]);
          split_tf32(dc[8], ahi[1], alo[1]);
          split_tf32(dc[4 * ldd], ahi[2], alo[2]);
          split_tf32(dc[4 * ldd + 8], ahi[3], alo[3]);
          const float* zr = zb + (size_t)s * 8 * ldz;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (nb < nbc) {
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(zr[nb * 8], bh0, bl0);
              split_tf32(zr[4 * ldz + nb * 8], bh1, bl1);
              mma3(accH[nb], ahi, alo, bh0, bh1, bl0, bl1, single);
            }
          }
        }
        if (!ACC) {
          float* row0 = dHpart + (size_t)(mt * 16 + g) * KG;
          float* row1 = row0 + (size_t)8 * KG;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (nb < nbc) {
              const int c0 = (nt0 + nb) * 8 + 2 * t;
              atomicAdd(reinterpret_cast<float2*>(row0 + c0), make_float2(accH[nb][0], accH[nb][1]));
              atomicAdd(reinterpret_cast<float2*>(row1 + c0), make_float2(accH[nb][2], accH[nb][3]));
            }
          }
        }
      }
      __syncthreads();  // all reads of Z done before U overwrites it
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 3);
    }

    if (want_dx) {
      // ---- U[r][c] = sum_f D[r][f] h[f][c]  -> Z
      for (int task = warp; task < ntasks_u; task += CFG::kWarps) {
        const int mt = task / ngroups_u, ng = task - mt * ngroups_u;
        const int nt0 = ng * NB;
        const int nbc = min(NB, NTu - nt0);
        float acc[NB][4];
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const float* da = Ds + (size_t)(mt * 16 + g) * ldd + t;
        const float4* hp = HP + (size_t)nt0 * 32 + lane;
#pragma unroll 4
        for (int s = 0; s < KSu; ++s) {
          uint32_t ahi[4], alo[4];
          split_tf32(da[s * 8], ahi[0], alo[0]);
          split_tf32(da[8 * ldd + s * 8], ahi[1], alo[1]);
          split_tf32(da[s * 8 + 4], ahi[2], alo[2]);
          split_tf32(da[8 * ldd + s * 8 + 4], ahi[3], alo[3]);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (nb < nbc) {
              const float4 b = hp[((size_t)s * NTu + nb) * 32];
              mma3(acc[nb], ahi, alo, __float_as_uint(b.x), __float_as_uint(b.y),
                   __float_as_uint(b.z), __float_as_uint(b.w), single);
            }
          }
        }
        float* zr0 = Zs + (size_t)(mt * 16 + g) * ldz + nt0 * 8 + 2 * t;
        float* zr1 = zr0 + (size_t)8 * ldz;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (nb < nbc) {
            *reinterpret_cast<float2*>(zr0 + nb * 8) = make_float2(acc[nb][0], acc[nb][1]);
            *reinterpret_cast<float2*>(zr1 + nb * 8) = make_float2(acc[nb][2], acc[nb][3]);
          }
        }
      }
      __syncthreads();
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 4);
      // ---- Horner: acc = U_{K-1}; acc = acc S^T + U_k  (in place in slot k)
      if constexpr (kFast) {
        if (has_col) {   // Horner in registers, dX column stored straight to global
          horner_from_column<CFG>(zcol, Zs, Ss + (size_t)jcol * N * N, jcol, gcol, ldz);
          float* dst = a.dX + ((size_t)(b0 + jcol) * G + gcol) * N;
          if ((N & 3) == 0 && a.vec_ok) {
#pragma unroll
            for (int q = 0; q < N / 4; ++q)
              reinterpret_cast<float4*>(dst)[q] = make_float4(zcol[4 * q], zcol[4 * q + 1], zcol[4 * q + 2], zcol[4 * q + 3]);
          } else {
#pragma unroll
            for (int n = 0; n < N; ++n) dst[n] = zcol[n];
          }
        }
      } else {
        for (int k = K - 2; k >= 0; --k) {
          if (p.use_lists) hop_tile_lists<CFG, true>(Zs, Ss, lst_r, p, rows_used, ldz, (k + 1) * G, k * G);
          else hop_tile<CFG, true>(Zs, Ss, p, rows_used, ldz, (k + 1) * G, k * G);
          __syncthreads();
        }
        store_dx_tile<CFG>(Zs, a.dX, p, b0, gcount, ldz, a.vec_ok);
      }
    }
    __syncthreads();
    if (tile == (int)blockIdx.x) GFC_STAMP(a, 5);
  }

  if (want_dh && ACC) {
    // each CTA owns one full partial; tasks cover the whole [F x KG] range.
    const int task = warp;
    if (task < ntasks_d) {
      const int mt = task / ngroups_d, ng = task - mt * ngroups_d;
      const int nt0 = ng * nbd;
      const int nbc = min(nbd, NTd - nt0);
      float* row0 = dHpart + (size_t)(mt * 16 + g) * KG;
      float* row1 = row0 + (size_t)8 * KG;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (nb < nbc) {
          const int c0 = (nt0 + nb) * 8 + 2 * t;
          *reinterpret_cast<float2*>(row0 + c0) = make_float2(accH[nb][0], accH[nb][1]);
          *reinterpret_cast<float2*>(row1 + c0) = make_float2(accH[nb][2], accH[nb][3]);
        }
      }
    }
  }
  if (want_db) {
    for (int f = tid; f < F; f += CFG::kThreads) a.dbp[(size_t)blockIdx.x * F + f] = dbs[f];
  }
  GFC_STAMP(a, 6);
  GFC_STAMP_NS(a, 9);
}

// ---------------------------------------------------------------------------
// launch helpers shared by the per-variant translation units
// ---------------------------------------------------------------------------
template <typename Kern>
static int launch_tile_kernel(Kern kern, const TileArgs& a, cudaStream_t st, const char* name) {
  if (a.p.smem_bytes > 48 * 1024)
    GFC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.p.smem_bytes));
  kern<<<a.p.grid, a.p.threads, a.p.smem_bytes, st>>>(a);
  GFC_LAUNCH_CHECK(name);
  return GFC_OK;
}

// Instantiates every (GSRC, HSMEM[, ACC]) combination a variant can be planned with.
#define GFC_DEFINE_TILE_LAUNCHERS(SUFFIX, CFG)                                                          \
  int tile_fwd_##SUFFIX(const TileArgs& a, int gsrc, cudaStream_t st) {                                 \
    if (gsrc == GSRC_DENSE) {                                                                           \
      if (a.p.h_smem) return launch_tile_kernel(tile_fwd_kernel<CFG, GSRC_DENSE, true>, a, st, "tile_fwd<dense,hsmem>"); \
      return launch_tile_kernel(tile_fwd_kernel<CFG, GSRC_DENSE, false>, a, st, "tile_fwd<dense,hglobal>");              \
    }                                                                                                   \
    if (a.p.h_smem) return launch_tile_kernel(tile_fwd_kernel<CFG, GSRC_POS, true>, a, st, "tile_fwd<pos,hsmem>");       \
    return launch_tile_kernel(tile_fwd_kernel<CFG, GSRC_POS, false>, a, st, "tile_fwd<pos,hglobal>");   \
  }                                                                                                     \
  int tile_bwd_##SUFFIX(const TileArgs& a, int gsrc, cudaStream_t st) {                                 \
    const int sel = (gsrc == GSRC_POS ? 4 : 0) | (a.p.h_smem ? 2 : 0) | (a.p.acc_regs ? 1 : 0);         \
    switch (sel) {                                                                                      \
      case 0: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_DENSE, false, false>, a, st, "tile_bwd<dense,hglobal,flush>"); \
      case 1: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_DENSE, false, true>, a, st, "tile_bwd<dense,hglobal,regs>");   \
      case 2: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_DENSE, true, false>, a, st, "tile_bwd<dense,hsmem,flush>");    \
      case 3: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_DENSE, true, true>, a, st, "tile_bwd<dense,hsmem,regs>");      \
      case 4: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_POS, false, false>, a, st, "tile_bwd<pos,hglobal,flush>");     \
      case 5: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_POS, false, true>, a, st, "tile_bwd<pos,hglobal,regs>");       \
      case 6: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_POS, true, false>, a, st, "tile_bwd<pos,hsmem,flush>");        \
      default: return launch_tile_kernel(tile_bwd_kernel<CFG, GSRC_POS, true, true>, a, st, "tile_bwd<pos,hsmem,regs>");         \
    }                                                                                                   \
  }

}  // namespace gfc
