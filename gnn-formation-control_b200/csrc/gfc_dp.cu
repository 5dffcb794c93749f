// gfc_dp.cu — data-parallel gradient exchange: ONE kernel that finishes the gradient of a rank (fixed-order sum of
// the per-CTA partials the backward kernels left behind) and all-reduces it over NVLink peer memory.
//
//   phase 1  each CTA owns a slice of the flat [dH | db] bucket: it sums the slice over this rank's partial buffers
//            and stores the result straight into slot `rank` of EVERY peer's exchange buffer (P2P stores through
//            NVSwitch; the local copy is an ordinary store), then releases a flag (epoch number) on every peer;
//   phase 2  the CTA waits until the flags of all ranks for its slice carry the current epoch and sums the `world`
//            slots in rank order (so every rank obtains bit-identical gradients), scales, writes the bucket.
// One-shot: 2 x (world-1) x n floats cross the links per rank, ~one NVLink round trip of latency; there is no
// grid-wide synchronisation (slices are independent) and no host involvement.  The exchange buffers are double
// buffered by epoch parity: a rank can only start epoch e+2 after every peer has finished reading epoch e,
// because it needs the peers' epoch e+1 flags to complete epoch e+1 first.
// The reference has no distributed code (single .to('cuda'), suhaas_agent.py:19); this replaces the
// reduce_parts_kernel + ncclAllReduce pair of the plain DP path.
#include "gfc_common.cuh"
#include "gfc_dp.cuh"
#include "gfc_tile.cuh"   // g_pdl

namespace gfc {

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct DpPeers {
  float* buf[GFC_DP_MAX_WORLD];          // exchange buffer of every rank: [2][world][n]
  unsigned int* sig[GFC_DP_MAX_WORLD];   // flags of every rank: [2][world][nblocks], then epochs [nblocks]
};

__global__ void __launch_bounds__(256)
reduce_allreduce_kernel(const float* __restrict__ pa, int npa, int na, const float* __restrict__ pb, int npb, int nb,
                        float* __restrict__ out, const DpPeers peers, int rank, int world, float scale) {
  __shared__ unsigned int s_epoch;
  // programmatic dependent launch: this grid may be scheduled while the backward kernel drains
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n = na + nb;
  const int nblocks = gridDim.x, c = blockIdx.x;
  const int per = (n + nblocks - 1) / nblocks;
  const int lo = c * per, hi = min(n, lo + per);
  unsigned int* my_sig = peers.sig[rank];
  unsigned int* epochs = my_sig + (size_t)2 * world * nblocks;
  if (threadIdx.x == 0) { s_epoch = epochs[c] + 1; epochs[c] = s_epoch; }
  __syncthreads();
  const unsigned int e = s_epoch;
  const int par = e & 1;
  // ---- phase 1: local fixed-order reduction of the slice, pushed to every rank ---------------------
  // 32 elements at a time: 8 partial-classes x 32 lanes (8 loads in flight per thread), then a fixed-order
  // combine through shared memory — the same order on every launch, so the gradients are deterministic
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, pg = threadIdx.x >> 5;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    float s = 0.f;
    if (i < hi) {
      const float* parts; int np, nn, ii;
      if (i < na) { parts = pa; np = npa; nn = na; ii = i; } else { parts = pb; np = npb; nn = nb; ii = i - na; }
      int p = pg;
      for (; p + 56 < np; p += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = parts[(size_t)(p + 8 * u) * nn + ii];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; p < np; p += 8) s += parts[(size_t)p * nn + ii];
    }
    red[pg][lane] = s;
    __syncthreads();
    if (pg == 0 && i < hi) {
      float t = red[0][lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += red[k][lane];
      for (int r = 0; r < world; ++r) peers.buf[r][((size_t)par * world + rank) * n + i] = t;
    }
    __syncthreads();
  }
  // (the barrier orders every thread's stores before thread t's system-scope release: cumulativity)
  __syncthreads();
  if ((int)threadIdx.x < world)
    st_release_sys(peers.sig[threadIdx.x] + ((size_t)par * world + rank) * nblocks + c, e);
  // ---- phase 2: wait for every rank's slice, sum in rank order ---------------------------------------
  if ((int)threadIdx.x < world) {
    const unsigned int* f = my_sig + ((size_t)par * world + threadIdx.x) * nblocks + c;
    while ((int)(ld_acquire_sys(f) - e) < 0) { __nanosleep(40); }
  }
  __syncthreads();
  const float* mine = peers.buf[rank] + (size_t)par * world * n;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcg(mine + (size_t)r * n + i);
    out[i] = s * scale;
  }
}

int dp_blocks(int n) {
  int nb = ceil_div(n, 32);
  if (nb > GFC_DP_MAX_BLOCKS) nb = GFC_DP_MAX_BLOCKS;
  if (nb < 1) nb = 1;
  return nb;
}

int launch_reduce_allreduce(const float* pa, int npa, int na, const float* pb, int npb, int nb, float* out,
                            const DpCtx& dp, cudaStream_t st) {
  GFC_REQUIRE(dp.world >= 1 && dp.world <= GFC_DP_MAX_WORLD && dp.rank >= 0 && dp.rank < dp.world, GFC_ERR_BAD_ARG,
              "dp: bad rank/world %d/%d", dp.rank, dp.world);
  GFC_REQUIRE(dp.peer_buf && dp.peer_sig && out, GFC_ERR_BAD_ARG, "dp: NULL exchange pointers");
  DpPeers peers;
  for (int r = 0; r < dp.world; ++r) {
    peers.buf[r] = static_cast<float*>(dp.peer_buf[r]);
    peers.sig[r] = static_cast<unsigned int*>(dp.peer_sig[r]);
    GFC_REQUIRE(peers.buf[r] && peers.sig[r], GFC_ERR_BAD_ARG, "dp: NULL peer pointer for rank %d", r);
  }
  const int n = na + nb;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(dp_blocks(n));
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  GFC_CUDA_TRY(cudaLaunchKernelEx(&cfg, reduce_allreduce_kernel, pa, npa, na, pb, npb, nb, out, peers, dp.rank,
                                  dp.world, dp.scale));
  GFC_LAUNCH_CHECK("reduce_allreduce_kernel");
  return GFC_OK;
}

}  // namespace gfc

using namespace gfc;

extern "C" size_t gfc_dp_exchange_bytes(int n, int world) {
  if (n <= 0 || world <= 0 || world > GFC_DP_MAX_WORLD) return 0;
  return align_up((size_t)2 * world * n * sizeof(float), 256);
}
extern "C" size_t gfc_dp_signal_bytes(int n, int world) {
  if (n <= 0 || world <= 0 || world > GFC_DP_MAX_WORLD) return 0;
  return align_up(((size_t)2 * world + 1) * dp_blocks(n) * sizeof(unsigned int), 256);
}
