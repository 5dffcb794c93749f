// gfc_csr_fused.cu — the whole forward of the CSR filter (kernel (d)) for one graph in ONE CTA:
//   y = act(sum_k z_k H_k + b),  z_0 = x,  z_k = z_{k-1} S   (BatchLSIGF, utils/graphUtils/graphML.py:2342-2366)
// The graph's current state z_k [N x G] lives in shared memory for the whole launch.  Per tap k the 32 warps
//   (1) contract their rows of z_k with H_k on the tensor cores (mma.sync m16n8k8 TF32, 3xTF32 split — the same
//       fragments as the tile kernels) and accumulate into y (read-modify-write of the L2-resident output tile;
//       bias joins at k = 0, the activation at k = K-1), then
//   (2) gather the next state from shared memory through the CSR lists (results held in registers until every
//       gather is done, then the state is overwritten in place).
// Nothing but x, the CSR lists and y touches global memory: no transposed copy of x, no workspace of diffusion
// states, no separate GEMM launch.  Used when a graph's state fits (N * G <= 32768 floats, e.g. cfg5: 1024 x 32).
#include "gfc_tile_kernels.cuh"
#include "gfc_generic.cuh"

namespace gfc {

constexpr int kFusedThreads = 1024, kFusedItems = 8;   // float4 gather results per thread

// One row of a hop: sum over the row's list entries i in [beg, end) of w_i * state[m_i] (a float4 = four channels).
// (Tried in round 2: reading the list four entries at a time, one group ahead of their use.  The index loads hit L1 and
// the kernels sit at the 64-register limit of a 1024-thread CTA: forward 269 -> 304 us, backward 453 -> 502 us. Dropped.)
__device__ __forceinline__ float4 gather_row(const float* __restrict__ src, int stride, const int32_t* __restrict__ ci,
                                             const float* __restrict__ vv, int beg, int end) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = beg; i < end; ++i) {
    const int m = ci[i];
    const float w = vv ? vv[i] : 1.f;
    const float4 z = *reinterpret_cast<const float4*>(src + (size_t)m * stride);
    acc.x = fmaf(w, z.x, acc.x);
    acc.y = fmaf(w, z.y, acc.y);
    acc.z = fmaf(w, z.z, acc.z);
    acc.w = fmaf(w, z.w, acc.w);
  }
  return acc;
}

// The same sum with the row's list read from SHARED memory as 16-bit node numbers, two per 32-bit load (same order of the
// additions, so the result is bit-identical).  The gathers are bound by L1 / shared-memory wavefronts (ncu, round 1): a
// global index load of a warp's four rows costs ~3 wavefronts per step next to the ~3.4 of the row reads themselves; a
// packed shared-memory word costs 1 per TWO steps.
__device__ __forceinline__ float4 gather_row_s16(const float* __restrict__ src, int stride, const uint16_t* __restrict__ cis,
                                                 const float* __restrict__ vv, int beg, int end) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = beg;
  if ((i & 1) && i < end) {
    const float w = vv ? vv[i] : 1.f;
    const float4 z = *reinterpret_cast<const float4*>(src + (size_t)cis[i] * stride);
    acc.x = fmaf(w, z.x, acc.x); acc.y = fmaf(w, z.y, acc.y); acc.z = fmaf(w, z.z, acc.z); acc.w = fmaf(w, z.w, acc.w);
    ++i;
  }
  for (; i + 1 < end; i += 2) {
    const uint32_t pr = *reinterpret_cast<const uint32_t*>(cis + i);
    const float w0 = vv ? vv[i] : 1.f, w1 = vv ? vv[i + 1] : 1.f;
    const float4 z0 = *reinterpret_cast<const float4*>(src + (size_t)(pr & 0xffffu) * stride);
    const float4 z1 = *reinterpret_cast<const float4*>(src + (size_t)(pr >> 16) * stride);
    acc.x = fmaf(w0, z0.x, acc.x); acc.y = fmaf(w0, z0.y, acc.y); acc.z = fmaf(w0, z0.z, acc.z); acc.w = fmaf(w0, z0.w, acc.w);
    acc.x = fmaf(w1, z1.x, acc.x); acc.y = fmaf(w1, z1.y, acc.y); acc.z = fmaf(w1, z1.z, acc.z); acc.w = fmaf(w1, z1.w, acc.w);
  }
  if (i < end) {
    const float w = vv ? vv[i] : 1.f;
    const float4 z = *reinterpret_cast<const float4*>(src + (size_t)cis[i] * stride);
    acc.x = fmaf(w, z.x, acc.x); acc.y = fmaf(w, z.y, acc.y); acc.z = fmaf(w, z.z, acc.z); acc.w = fmaf(w, z.w, acc.w);
  }
  return acc;
}
// idx / d and idx % d with a shift when d is a power of two (sh >= 0): the channel-group and node counts are run-time
// values, and the emulated 32-bit division was ~10 % of the forward kernel's instructions (ncu source page, round 2)
__device__ __forceinline__ int pow2_shift(int d) { return (d & (d - 1)) == 0 ? __ffs(d) - 1 : -1; }
__device__ __forceinline__ void divmod(int idx, int d, int sh, int& q, int& r) {
  if (sh >= 0) { q = idx >> sh; r = idx & (d - 1); } else { q = idx / d; r = idx - q * d; }
}
// first `icap` list entries of the graph -> shared memory as 16-bit node numbers; returns how many were staged
__device__ __forceinline__ int stage_indices(uint16_t* cis, const int32_t* __restrict__ ci, int nnz, long long nnz_stride, int icap) {
  int nst = nnz < icap ? nnz : icap;
  if ((long long)nst > nnz_stride) nst = (int)nnz_stride;
  // four entries per load (the lists of a graph start 16-byte aligned: nnz_stride is a multiple of 4), all loads of a
  // thread independent: the scalar loop cost one L2 round trip per 1024 entries (5 % of the forward kernel's samples)
  if ((reinterpret_cast<uintptr_t>(ci) & 15) == 0) {
    const int4* c4 = reinterpret_cast<const int4*>(ci);
    uint2* o2 = reinterpret_cast<uint2*>(cis);
#pragma unroll 4
    for (int i = threadIdx.x; i < (nst >> 2); i += blockDim.x) {
      const int4 v = __ldg(c4 + i);
      o2[i] = make_uint2((uint32_t)v.x | ((uint32_t)v.y << 16), (uint32_t)v.z | ((uint32_t)v.w << 16));
    }
    for (int i = (nst & ~3) + threadIdx.x; i < nst; i += blockDim.x) cis[i] = (uint16_t)ci[i];
  } else {
    for (int i = threadIdx.x; i < nst; i += blockDim.x) cis[i] = (uint16_t)ci[i];
  }
  return nst;
}

template <int NT>   // F / 8 output column tiles, compile time: the accumulators must stay in registers
__global__ void __launch_bounds__(kFusedThreads, 1)
csr_fwd_fused_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ colidx, const float* __restrict__ vals, long long nnz_stride,
                     const float* __restrict__ h, const float* __restrict__ bias, float* __restrict__ y,
                     int N, int G, int F, int K, int act, float slope, int single, int icap) {
  extern __shared__ __align__(16) float smem[];
  const int GS = G + 4;                                   // row stride of the state: conflict-free A fragments
  const int MT = (N + 15) >> 4, KS = G >> 3, KG = K * G;
  float* zs = smem;                                       // [MT*16][GS]
  float4* Hs = reinterpret_cast<float4*>(smem + (size_t)MT * 16 * GS);   // packed taps, B-fragment order
  uint16_t* cis = reinterpret_cast<uint16_t*>(smem + (size_t)MT * 16 * GS + (size_t)KG * F * 2);   // [icap] staged list entries
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int G4 = G >> 2, total = N * G4;
  const int g4sh = pow2_shift(G4), nsh = pow2_shift(N);
  const int32_t* rp = rowptr + (size_t)b * (N + 1);
  const int32_t* ci = colidx + (size_t)b * nnz_stride;
  const float* vv = vals ? vals + (size_t)b * nnz_stride : nullptr;
  float* yb = y + (size_t)b * N * F;

  // ---- taps -> B fragments (hi / lo); x [G][N] -> state [N][GS] (transposed on the fly); pad rows = 0 --------
  for (int q = tid; q < (KG * F) >> 1; q += kFusedThreads) Hs[q] = pack_one(h, F, KG, 0, q);
  {
    const float* xb = x + (size_t)b * G * N;
    for (int idx = tid; idx < G * N; idx += kFusedThreads) {
      int gg, n;
      divmod(idx, N, nsh, gg, n);
      zs[(size_t)n * GS + gg] = __ldg(xb + idx);
    }
    for (int idx = N * GS + tid; idx < MT * 16 * GS; idx += kFusedThreads) zs[idx] = 0.f;
  }
  const int nst = (K > 1 && icap > 0) ? stage_indices(cis, ci, rp[N], nnz_stride, icap) : 0;
  __syncthreads();

  for (int k = 0; k < K; ++k) {
    // ---- (1) y[rows of this warp] (+)= z_k H_k ------------------------------------------------------------
    for (int mt = warp; mt < MT; mt += kFusedThreads / 32) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      float* y0 = yb + (size_t)r0 * F + 2 * t;
      float* y1 = yb + (size_t)r1 * F + 2 * t;
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        {
          if (k == 0) {
            const float b0 = bias ? __ldg(bias + nt * 8 + 2 * t) : 0.f, b1 = bias ? __ldg(bias + nt * 8 + 2 * t + 1) : 0.f;
            acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
          } else {
            const float2 u = r0 < N ? *reinterpret_cast<const float2*>(y0 + nt * 8) : make_float2(0.f, 0.f);
            const float2 v = r1 < N ? *reinterpret_cast<const float2*>(y1 + nt * 8) : make_float2(0.f, 0.f);
            acc[nt][0] = u.x; acc[nt][1] = u.y; acc[nt][2] = v.x; acc[nt][3] = v.y;
          }
        }
      }
      const float* za = zs + (size_t)r0 * GS + t;
      const float4* hp = Hs + (size_t)(k * KS) * NT * 32 + lane;
      for (int s = 0; s < KS; ++s) {
        uint32_t ahi[4], alo[4];
        split_tf32(za[s * 8], ahi[0], alo[0]);
        split_tf32(za[8 * GS + s * 8], ahi[1], alo[1]);
        split_tf32(za[s * 8 + 4], ahi[2], alo[2]);
        split_tf32(za[8 * GS + s * 8 + 4], ahi[3], alo[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          {
            const float4 bf = hp[((size_t)s * NT + nt) * 32];
            mma3(acc[nt], ahi, alo, __float_as_uint(bf.x), __float_as_uint(bf.y), __float_as_uint(bf.z),
                 __float_as_uint(bf.w), single != 0);
          }
        }
      }
      const bool last = k == K - 1;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        {
          if (last) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = apply_act(acc[nt][i], act, slope);
          }
          if (r0 < N) *reinterpret_cast<float2*>(y0 + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
          if (r1 < N) *reinterpret_cast<float2*>(y1 + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
    }
    if (k == K - 1) break;
    // ---- (2) z_{k+1}[n] = sum_m z_k[m] S[m][n]: the lists of node n name the m (transposed use is the caller's) --
    float4 nxt[kFusedItems];
#pragma unroll
    for (int it = 0; it < kFusedItems; ++it) {
      const int idx = tid + it * kFusedThreads;
      nxt[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) {
        int n, g4;
        divmod(idx, G4, g4sh, n, g4);
        const int beg = rp[n], end = rp[n + 1];
        nxt[it] = end <= nst ? gather_row_s16(zs + g4 * 4, GS, cis, vv, beg, end) : gather_row(zs + g4 * 4, GS, ci, vv, beg, end);
      }
    }
    __syncthreads();                                      // every tap MMA and every gather of z_k is done
#pragma unroll
    for (int it = 0; it < kFusedItems; ++it) {
      const int idx = tid + it * kFusedThreads;
      if (idx < total) {
        int n, g4;
        divmod(idx, G4, g4sh, n, g4);
        *reinterpret_cast<float4*>(zs + (size_t)n * GS + g4 * 4) = nxt[it];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, same structure (closed form of the autograd graph, no recomputation of the forward states):
//   V_0 = dY o act'(y),  V_k = V_{k-1} S^T  (transposed lists),
//   dX = sum_k V_k H_k^T,   dH_k = V_k^T X,   db = column sums of V_0.
// The state V_k [N x F] lives in shared memory.  Per tap: (a) dX (+)= V_k H_k^T on the tensor cores, accumulated by
// read-modify-write of the L2-resident dX tile (feature-major [G][N], 32-byte segments); (b) dH_k: the node
// dimension is split over the 32 warps (3xTF32 MMAs with the X fragments read from x in L2), partial tiles meet in
// a shared-memory accumulator through shared atomics; (c) the next state is gathered through the transposed
// lists.  One partial [F x K*G] (+ db) per graph leaves the CTA; reduce_parts_kernel sums them in a fixed order.
template <int NTG>   // G / 8
__global__ void __launch_bounds__(kFusedThreads, 1)
csr_bwd_fused_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr_t,
                     const int32_t* __restrict__ colidx_t, const float* __restrict__ vals_t, long long nnz_stride,
                     const float* __restrict__ h, const float* __restrict__ yout, const float* __restrict__ dY,
                     float* __restrict__ dX, float* __restrict__ dHp, float* __restrict__ dbp,
                     int N, int G, int F, int K, int act, float slope, int single, int icap) {
  extern __shared__ __align__(16) float smem[];
  const int FS = F + 4;
  const int MT = (N + 15) >> 4, KSF = F >> 3, KG = K * G, NTC = KG >> 3, MTF = F >> 4;
  float* vs = smem;                                                        // [MT*16][FS]
  float4* Hs = reinterpret_cast<float4*>(smem + (size_t)MT * 16 * FS);     // taps, B fragments of H_k^T (for_bwd)
  float* dHs = smem + (size_t)MT * 16 * FS + (size_t)KG * F * 2;           // [F][KG] running dH of this graph
  float* dbs = dHs + (size_t)F * KG;                                       // [F]
  float* dbw = dbs + F;                                                    // [32 warps][F] per-warp column sums of V_0
  uint16_t* cis = reinterpret_cast<uint16_t*>(dbw + (kFusedThreads / 32) * F);   // [icap] staged (transposed) list entries
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int F4 = F >> 2, total = N * F4;
  const int f4sh = pow2_shift(F4);
  const int32_t* rp = rowptr_t + (size_t)b * (N + 1);
  const int32_t* ci = colidx_t + (size_t)b * nnz_stride;
  const float* vv = vals_t ? vals_t + (size_t)b * nnz_stride : nullptr;
  const float* xb = x ? x + (size_t)b * G * N : nullptr;
  float* dxb = dX ? dX + (size_t)b * G * N : nullptr;
  const bool want_dh = dHp != nullptr && xb != nullptr, want_db = dbp != nullptr;

  // ---- setup: taps, zeroed accumulators, V_0 = dY o act'(y) (node-major, coalesced), db ---------------------
  if (dxb)
    for (int q = tid; q < (KG * F) >> 1; q += kFusedThreads) Hs[q] = pack_one(h, F, KG, 1, q);
  for (int i = tid; i < F * KG + F; i += kFusedThreads) dHs[i] = 0.f;
  for (int idx = N * FS + tid; idx < MT * 16 * FS; idx += kFusedThreads) vs[idx] = 0.f;
  const int nst = (K > 1 && icap > 0) ? stage_indices(cis, ci, rp[N], nnz_stride, icap) : 0;
  __syncthreads();
  {
    const float4* d4 = reinterpret_cast<const float4*>(dY + (size_t)b * N * F);
    const float4* y4 = (act != GFC_ACT_NONE) ? reinterpret_cast<const float4*>(yout + (size_t)b * N * F) : nullptr;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int idx = tid; idx < total; idx += kFusedThreads) {     // idx % F4 is the same for every pass of a thread
      int n, f4;
      divmod(idx, F4, f4sh, n, f4);
      float4 v = __ldg(d4 + idx);
      if (y4) {
        const float4 yo = __ldg(y4 + idx);
        v.x = act_grad(v.x, yo.x, act, slope); v.y = act_grad(v.y, yo.y, act, slope);
        v.z = act_grad(v.z, yo.z, act, slope); v.w = act_grad(v.w, yo.w, act, slope);
      }
      *reinterpret_cast<float4*>(vs + (size_t)n * FS + f4 * 4) = v;
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
    if (want_db) {
      // deterministic: lanes with the same column group (lane % F4) meet by shuffles, the 32 per-warp sums in fixed order
      for (int o = F4; o < 32; o <<= 1) {
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
      }
      if (lane < F4) *reinterpret_cast<float4*>(dbw + (size_t)warp * F + lane * 4) = cs;
    }
  }
  __syncthreads();
  if (want_db && tid < F) {
    float sum = 0.f;
    for (int wq = 0; wq < kFusedThreads / 32; ++wq) sum += dbw[(size_t)wq * F + tid];
    dbs[tid] = sum;
  }

  for (int k = 0; k < K; ++k) {
    // ---- (a) dX[rows of this warp][g] (+)= sum_f V_k[row][f] h[f][k*G + g] ----------------------------------
    if (dxb) {
      for (int mt = warp; mt < MT; mt += kFusedThreads / 32) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        float acc[NTG][4];
#pragma unroll
        for (int nt = 0; nt < NTG; ++nt) {
          float* c0 = dxb + (size_t)(nt * 8 + 2 * t) * N;
          if (k == 0) {
            acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
          } else {
            acc[nt][0] = r0 < N ? c0[r0] : 0.f; acc[nt][1] = r0 < N ? c0[N + r0] : 0.f;
            acc[nt][2] = r1 < N ? c0[r1] : 0.f; acc[nt][3] = r1 < N ? c0[N + r1] : 0.f;
          }
        }
        const float* va = vs + (size_t)r0 * FS + t;
        const float4* hp = Hs + (size_t)(k * NTG) * 32 + lane;
        for (int s = 0; s < KSF; ++s) {
          uint32_t ahi[4], alo[4];
          split_tf32(va[s * 8], ahi[0], alo[0]);
          split_tf32(va[8 * FS + s * 8], ahi[1], alo[1]);
          split_tf32(va[s * 8 + 4], ahi[2], alo[2]);
          split_tf32(va[8 * FS + s * 8 + 4], ahi[3], alo[3]);
#pragma unroll
          for (int nt = 0; nt < NTG; ++nt) {
            const float4 bf = hp[((size_t)s * NTC + nt) * 32];
            mma3(acc[nt], ahi, alo, __float_as_uint(bf.x), __float_as_uint(bf.y), __float_as_uint(bf.z),
                 __float_as_uint(bf.w), single != 0);
          }
        }
#pragma unroll
        for (int nt = 0; nt < NTG; ++nt) {
          float* c0 = dxb + (size_t)(nt * 8 + 2 * t) * N;
          if (r0 < N) { c0[r0] = acc[nt][0]; c0[N + r0] = acc[nt][1]; }
          if (r1 < N) { c0[r1] = acc[nt][2]; c0[N + r1] = acc[nt][3]; }
        }
      }
    }
    // ---- (b) dH[f][k*G + g] += sum_n V_k[n][f] x[g][n]: a warp owns ONE 16 x 8 output tile and a share of the
    //          node dimension (few partial sums per address meet in the shared accumulator) ---------------------
    if (want_dh) {
      const int ntile = MTF * NTG, ngrp = (kFusedThreads / 32) / ntile;
      const int tile = warp % ntile, grp = warp / ntile;
      const int mtf = tile / NTG, nt = tile - mtf * NTG;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* xc = xb + (size_t)(nt * 8 + g) * N;                // B[k = node][n = g] = x[g][node]
      // 8 nodes per step; the X fragments of FOUR steps are requested before the first is used: with one step at a time
      // every step waited an L2 round trip (24 % of the kernel's stall samples sat on the first use of these loads)
      for (int s4 = grp; s4 < 2 * MT; s4 += 4 * ngrp) {
        float xv[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n0 = (s4 + u * ngrp) * 8 + t, n1 = n0 + 4;         // s >= 2 MT gives n0 >= N: zero
          xv[u][0] = n0 < N ? __ldg(xc + n0) : 0.f;
          xv[u][1] = n1 < N ? __ldg(xc + n1) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s = s4 + u * ngrp;
          if (s < 2 * MT) {                                             // warp-uniform
            const int n0 = s * 8 + t;
            const float* va = vs + (size_t)n0 * FS + mtf * 16 + g;      // A[m = f][k = node] = V[node][f]
            uint32_t ahi[4], alo[4], bh0, bl0, bh1, bl1;
            split_tf32(xv[u][0], bh0, bl0);
            split_tf32(xv[u][1], bh1, bl1);
            split_tf32(va[0], ahi[0], alo[0]);
            split_tf32(va[8], ahi[1], alo[1]);
            split_tf32(va[4 * FS], ahi[2], alo[2]);
            split_tf32(va[4 * FS + 8], ahi[3], alo[3]);
            mma3(acc, ahi, alo, bh0, bh1, bl0, bl1, single != 0);
          }
        }
      }
      float* row0 = dHs + (size_t)(mtf * 16 + g) * KG + k * G + nt * 8 + 2 * t;
      float* row1 = row0 + (size_t)8 * KG;
      // the node groups add their partial tiles one after the other: fixed order, bitwise reproducible gradients
      for (int gi = 0; gi < ngrp; ++gi) {
        if (grp == gi) { row0[0] += acc[0]; row0[1] += acc[1]; row1[0] += acc[2]; row1[1] += acc[3]; }
        __syncthreads();
      }
    }
    if (k == K - 1) break;
    // ---- (c) V_{k+1}[n] = sum_m S[n][m] V_k[m]: transposed lists ----------------------------------------------
    float4 nxt[kFusedItems];
#pragma unroll
    for (int it = 0; it < kFusedItems; ++it) {
      const int idx = tid + it * kFusedThreads;
      nxt[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) {
        int n, f4;
        divmod(idx, F4, f4sh, n, f4);
        const int beg = rp[n], end = rp[n + 1];
        nxt[it] = end <= nst ? gather_row_s16(vs + f4 * 4, FS, cis, vv, beg, end) : gather_row(vs + f4 * 4, FS, ci, vv, beg, end);
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kFusedItems; ++it) {
      const int idx = tid + it * kFusedThreads;
      if (idx < total) {
        int n, f4;
        divmod(idx, F4, f4sh, n, f4);
        *reinterpret_cast<float4*>(vs + (size_t)n * FS + f4 * 4) = nxt[it];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (dHp)
    for (int i = tid; i < F * KG; i += kFusedThreads) dHp[(size_t)b * F * KG + i] = dHs[i];
  if (want_db)
    for (int i = tid; i < F; i += kFusedThreads) dbp[(size_t)b * F + i] = dbs[i];
}

int g_csr_stage_idx = 1;   // gfc_set_option(GFC_OPT_CSR_STAGE_IDX)
// list entries (16-bit) that fit behind the kernel's own shared memory; 0 = no staging
static int index_capacity(size_t bytes, int N) {
  DeviceInfo di;
  if (!g_csr_stage_idx || N > 65535 || get_device_info(&di)) return 0;
  const long long room = (long long)di.smem_optin - 1024 - (long long)bytes;
  return room < 16 ? 0 : (int)((room / 2) & ~7LL);
}

bool csr_bwd_fused_supported(int N, int G, int F, int K, size_t* smem_bytes) {
  if (N < 1 || K < 1 || !(G == 16 || G == 32) || (F & 15) || F > 64) return false;
  if ((kFusedThreads % (F >> 2)) != 0) return false;
  if ((kFusedThreads / 32) % ((F >> 4) * (G >> 3)) != 0) return false;   // warps = output tiles x node groups
  if ((long long)N * F > (long long)kFusedThreads * kFusedItems * 4) return false;
  const size_t MT = (size_t)(N + 15) >> 4;
  const size_t bytes = (MT * 16 * (F + 4) + (size_t)K * G * F * 2 + (size_t)F * K * G + F + (size_t)(kFusedThreads / 32) * F) * sizeof(float);
  DeviceInfo di;
  if (get_device_info(&di)) return false;
  if (bytes + 1024 > (size_t)di.smem_optin) return false;
  if (smem_bytes) *smem_bytes = bytes;
  return true;
}

int launch_csr_bwd_fused(const float* x, const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                         long long nnz_stride, const float* h, const float* yout, const float* dY, float* dX,
                         float* dHp, float* dbp, int B, int N, int G, int F, int K, int act, float slope, int single,
                         cudaStream_t st) {
  size_t smem = 0;
  if (!csr_bwd_fused_supported(N, G, F, K, &smem)) return GFC_ERR_UNSUPPORTED;
  const int icap = index_capacity(smem, N);
  smem += (size_t)icap * 2;
  if (G == 32) {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_bwd_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_bwd_fused_kernel<4><<<B, kFusedThreads, smem, st>>>(x, rowptr_t, colidx_t, vals_t, nnz_stride, h, yout, dY, dX,
                                                            dHp, dbp, N, G, F, K, act, slope, single, icap);
  } else {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_bwd_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_bwd_fused_kernel<2><<<B, kFusedThreads, smem, st>>>(x, rowptr_t, colidx_t, vals_t, nnz_stride, h, yout, dY, dX,
                                                            dHp, dbp, N, G, F, K, act, slope, single, icap);
  }
  GFC_LAUNCH_CHECK("csr_bwd_fused_kernel");
  return GFC_OK;
}

bool csr_fwd_fused_supported(int N, int G, int F, int K, size_t* smem_bytes) {
  if (N < 1 || K < 1 || (G & 7) || !(F == 16 || F == 32)) return false;   // instantiated output widths
  if ((long long)N * G > (long long)kFusedThreads * kFusedItems * 4) return false;
  const size_t MT = (size_t)(N + 15) >> 4;
  const size_t bytes = (MT * 16 * (G + 4) + (size_t)K * G * F * 2) * sizeof(float);
  DeviceInfo di;
  if (get_device_info(&di)) return false;
  if (bytes + 1024 > (size_t)di.smem_optin) return false;
  if (smem_bytes) *smem_bytes = bytes;
  return true;
}

int launch_csr_fwd_fused(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         long long nnz_stride, const float* h, const float* bias, float* y, int B, int N, int G,
                         int F, int K, int act, float slope, int single, cudaStream_t st) {
  size_t smem = 0;
  if (!csr_fwd_fused_supported(N, G, F, K, &smem)) return GFC_ERR_UNSUPPORTED;
  const int icap = index_capacity(smem, N);
  smem += (size_t)icap * 2;
  if (F == 32) {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_fwd_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_fwd_fused_kernel<4><<<B, kFusedThreads, smem, st>>>(x, rowptr, colidx, vals, nnz_stride, h, bias, y, N, G, F, K,
                                                            act, slope, single, icap);
  } else {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_fwd_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_fwd_fused_kernel<2><<<B, kFusedThreads, smem, st>>>(x, rowptr, colidx, vals, nnz_stride, h, bias, y, N, G, F, K,
                                                            act, slope, single, icap);
  }
  GFC_LAUNCH_CHECK("csr_fwd_fused_kernel");
  return GFC_OK;
}

}  // namespace gfc
