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

template <int NT>   // F / 8 output column tiles, compile time: the accumulators must stay in registers
__global__ void __launch_bounds__(kFusedThreads, 1)
csr_fwd_fused_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ colidx, const float* __restrict__ vals, long long nnz_stride,
                     const float* __restrict__ h, const float* __restrict__ bias, float* __restrict__ y,
                     int N, int G, int F, int K, int act, float slope, int single) {
  extern __shared__ __align__(16) float smem[];
  const int GS = G + 4;                                   // row stride of the state: conflict-free A fragments
  const int MT = (N + 15) >> 4, KS = G >> 3, KG = K * G;
  float* zs = smem;                                       // [MT*16][GS]
  float4* Hs = reinterpret_cast<float4*>(smem + (size_t)MT * 16 * GS);   // packed taps, B-fragment order
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int G4 = G >> 2, total = N * G4;
  const int32_t* rp = rowptr + (size_t)b * (N + 1);
  const int32_t* ci = colidx + (size_t)b * nnz_stride;
  const float* vv = vals ? vals + (size_t)b * nnz_stride : nullptr;
  float* yb = y + (size_t)b * N * F;

  // ---- taps -> B fragments (hi / lo); x [G][N] -> state [N][GS] (transposed on the fly); pad rows = 0 --------
  for (int q = tid; q < (KG * F) >> 1; q += kFusedThreads) Hs[q] = pack_one(h, F, KG, 0, q);
  {
    const float* xb = x + (size_t)b * G * N;
    for (int idx = tid; idx < G * N; idx += kFusedThreads) {
      const int gg = idx / N, n = idx - gg * N;
      zs[(size_t)n * GS + gg] = __ldg(xb + idx);
    }
    for (int idx = N * GS + tid; idx < MT * 16 * GS; idx += kFusedThreads) zs[idx] = 0.f;
  }
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
        const int n = idx / G4, g4 = idx - n * G4;
        const int beg = rp[n], end = rp[n + 1];
        const float* src = zs + g4 * 4;
        for (int i = beg; i < end; ++i) {
          const int m = ci[i];
          const float w = vv ? vv[i] : 1.f;
          const float4 z = *reinterpret_cast<const float4*>(src + (size_t)m * GS);
          nxt[it].x = fmaf(w, z.x, nxt[it].x);
          nxt[it].y = fmaf(w, z.y, nxt[it].y);
          nxt[it].z = fmaf(w, z.z, nxt[it].z);
          nxt[it].w = fmaf(w, z.w, nxt[it].w);
        }
      }
    }
    __syncthreads();                                      // every tap MMA and every gather of z_k is done
#pragma unroll
    for (int it = 0; it < kFusedItems; ++it) {
      const int idx = tid + it * kFusedThreads;
      if (idx < total) {
        const int n = idx / G4, g4 = idx - n * G4;
        *reinterpret_cast<float4*>(zs + (size_t)n * GS + g4 * 4) = nxt[it];
      }
    }
    __syncthreads();
  }
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
  if (F == 32) {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_fwd_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_fwd_fused_kernel<4><<<B, kFusedThreads, smem, st>>>(x, rowptr, colidx, vals, nnz_stride, h, bias, y, N, G, F, K,
                                                            act, slope, single);
  } else {
    GFC_CUDA_TRY(cudaFuncSetAttribute(csr_fwd_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    csr_fwd_fused_kernel<2><<<B, kFusedThreads, smem, st>>>(x, rowptr, colidx, vals, nnz_stride, h, bias, y, N, G, F, K,
                                                            act, slope, single);
  }
  GFC_LAUNCH_CHECK("csr_fwd_fused_kernel");
  return GFC_OK;
}

}  // namespace gfc
